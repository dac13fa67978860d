// K1 instantiations for target family "dgauss2d" (see hmc_kernel.cuh); compiled once per math mode.
#define GM_TAG TagDiffGauss2D
#define GM_FN dgauss2d
#define GM_FIT 0
#define GM_2D 1
#include "k_target.inc"
