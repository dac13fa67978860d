// K2 (Metropolis–Hastings) — exact math mode (compiled with --fmad=false); see mh_kernel.cuh.
#define GM_EXACT 1
#include "mh_kernel.cuh"
