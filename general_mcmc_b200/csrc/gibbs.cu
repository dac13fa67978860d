// gibbs.cu — built-in conditionals of the Gibbs sweep kernel (gibbs_kernel.cuh), compiled with --fmad=false.
#include "gibbs_kernel.cuh"

namespace gm {

int gibbs_max_dim() { return 8; }

cudaError_t launch_gibbs(const GibbsLaunch& L, cudaStream_t st) {
  if (L.kind == 1) return gibbs_launch<CondMixtureXZ>(L, st);
  if (L.kind != 0) return cudaErrorInvalidValue;
  switch (L.dim) {
    case 1: return gibbs_launch<CondConstant<1>>(L, st);
    case 2: return gibbs_launch<CondConstant<2>>(L, st);
    case 3: return gibbs_launch<CondConstant<3>>(L, st);
    case 4: return gibbs_launch<CondConstant<4>>(L, st);
    case 5: return gibbs_launch<CondConstant<5>>(L, st);
    case 6: return gibbs_launch<CondConstant<6>>(L, st);
    case 7: return gibbs_launch<CondConstant<7>>(L, st);
    case 8: return gibbs_launch<CondConstant<8>>(L, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace gm
