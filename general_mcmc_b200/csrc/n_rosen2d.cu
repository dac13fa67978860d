// K5 (NUTS) instantiations for target family "rosen2d" (see nuts_kernel.cuh); compiled once per math mode.
#define GM_TAG TagRosenbrock2D
#define GM_FN rosen2d
#define GM_2D 1
#include "nuts_target.inc"
