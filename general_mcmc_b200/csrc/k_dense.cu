// K1 instantiations for target family "dense" (see hmc_kernel.cuh); compiled once per math mode.
#define GM_TAG TagDenseGauss
#define GM_FN dense
#define GM_FIT 0
#define GM_2D 0
#include "k_target.inc"
