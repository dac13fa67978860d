// K1 instantiations for target family "rosen" (see hmc_kernel.cuh); compiled once per math mode.
#define GM_TAG TagRosenbrockND
#define GM_FN rosen
#define GM_FIT 1
#define GM_2D 0
#include "k_target.inc"
