// K5 (NUTS) instantiations for target family "dense" (see nuts_kernel.cuh); compiled once per math mode.
#define GM_TAG TagDenseGauss
#define GM_FN dense
#define GM_2D 0
#include "nuts_target.inc"
