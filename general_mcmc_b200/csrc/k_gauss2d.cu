// K1 instantiations for target family "gauss2d" (see hmc_kernel.cuh); compiled once per math mode.
#define GM_TAG TagGauss2D
#define GM_FN gauss2d
#define GM_FIT 0
#define GM_2D 1
#include "k_target.inc"
