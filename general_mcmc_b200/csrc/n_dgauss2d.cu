// K5 (NUTS) instantiations for target family "dgauss2d" (see nuts_kernel.cuh); compiled once per math mode.
#define GM_TAG TagDiffGauss2D
#define GM_FN dgauss2d
#define GM_2D 1
#include "nuts_target.inc"
