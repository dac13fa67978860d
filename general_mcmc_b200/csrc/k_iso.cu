// K1 instantiations for target family "iso" (see hmc_kernel.cuh); compiled once per math mode.
#define GM_TAG TagIsoGauss
#define GM_FN iso
#define GM_FIT 0
#define GM_2D 0
#include "k_target.inc"
