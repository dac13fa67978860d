// K5 (NUTS) instantiations for target family "mix" (see nuts_kernel.cuh); compiled once per math mode.
#define GM_TAG TagMixture
#define GM_FN mix
#define GM_2D 0
#include "nuts_target.inc"
