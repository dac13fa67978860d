// K5 (NUTS) instantiations for target family "iso" (see nuts_kernel.cuh); compiled once per math mode.
#define GM_TAG TagIsoGauss
#define GM_FN iso
#define GM_2D 0
#include "nuts_target.inc"
