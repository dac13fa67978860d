// K1 instantiations for target family "rosen2d" (see hmc_kernel.cuh); compiled once per math mode.
#define GM_TAG TagRosenbrock2D
#define GM_FN rosen2d
#define GM_FIT 0
#define GM_2D 1
#include "k_target.inc"
