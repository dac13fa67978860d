// K5 (NUTS) instantiations for target family "rosen" (see nuts_kernel.cuh); compiled once per math mode.
#define GM_TAG TagRosenbrockND
#define GM_FN rosen
#define GM_2D 0
#include "nuts_target.inc"
