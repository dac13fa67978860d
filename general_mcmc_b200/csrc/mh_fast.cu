// K2 (Metropolis–Hastings) — fast math mode; see mh_kernel.cuh.
#define GM_EXACT 0
#include "mh_kernel.cuh"
