// K1 instantiations for target family "mix" (see hmc_kernel.cuh); compiled once per math mode.
#define GM_TAG TagMixture
#define GM_FN mix
#define GM_FIT 0
#define GM_2D 0
#include "k_target.inc"
