// dispatch.cu — routes type-erased launch descriptors to the per-target translation units and
// picks the (elements-per-lane, lanes-per-chain) decomposition.
#include "kernels.h"

#include <cstdlib>

namespace gm {

#define GM_DECL(fn)                                                        \
  cudaError_t launch_hmc_##fn##_fast(const HmcLaunch&, cudaStream_t);      \
  cudaError_t launch_hmc_##fn##_exact(const HmcLaunch&, cudaStream_t);     \
  cudaError_t launch_eval_##fn##_fast(const EvalLaunch&, cudaStream_t);    \
  cudaError_t launch_eval_##fn##_exact(const EvalLaunch&, cudaStream_t);
GM_DECL(rosen) GM_DECL(iso) GM_DECL(dense) GM_DECL(mix) GM_DECL(rosen2d) GM_DECL(dgauss2d) GM_DECL(gauss2d)
#undef GM_DECL

#define GM_DECL_N(fn)                                                      \
  cudaError_t launch_nuts_##fn##_fast(const NutsLaunch&, cudaStream_t);    \
  cudaError_t launch_nuts_##fn##_exact(const NutsLaunch&, cudaStream_t);
GM_DECL_N(rosen) GM_DECL_N(iso) GM_DECL_N(dense) GM_DECL_N(mix) GM_DECL_N(rosen2d) GM_DECL_N(dgauss2d)
#undef GM_DECL_N

#define GM_ROUTE_N(mode, L, st)                           \
  switch ((L).tgt.kind) {                                 \
    case 0: return launch_nuts_iso_##mode(L, st);         \
    case 2: return launch_nuts_dgauss2d_##mode(L, st);    \
    case 3: return launch_nuts_dense_##mode(L, st);       \
    case 4: return launch_nuts_rosen2d_##mode(L, st);     \
    case 5: return launch_nuts_rosen_##mode(L, st);       \
    case 6: return launch_nuts_mix_##mode(L, st);         \
  }                                                       \
  return cudaErrorInvalidValue;

#define GM_ROUTE(prefix, mode, L, st)                     \
  switch ((L).tgt.kind) {                                 \
    case 0: return prefix##iso_##mode(L, st);             \
    case 1: return prefix##gauss2d_##mode(L, st);         \
    case 2: return prefix##dgauss2d_##mode(L, st);        \
    case 3: return prefix##dense_##mode(L, st);           \
    case 4: return prefix##rosen2d_##mode(L, st);         \
    case 5: return prefix##rosen_##mode(L, st);           \
    case 6: return prefix##mix_##mode(L, st);             \
  }                                                       \
  return cudaErrorInvalidValue;

cudaError_t launch_hmc_fast(const HmcLaunch& L, cudaStream_t st) { GM_ROUTE(launch_hmc_, fast, L, st) }
cudaError_t launch_hmc_exact(const HmcLaunch& L, cudaStream_t st) { GM_ROUTE(launch_hmc_, exact, L, st) }
cudaError_t launch_eval_fast(const EvalLaunch& E, cudaStream_t st) { GM_ROUTE(launch_eval_, fast, E, st) }
cudaError_t launch_eval_exact(const EvalLaunch& E, cudaStream_t st) { GM_ROUTE(launch_eval_, exact, E, st) }

cudaError_t launch_nuts_fast(const NutsLaunch& L, cudaStream_t st) { GM_ROUTE_N(fast, L, st) }
cudaError_t launch_nuts_exact(const NutsLaunch& L, cudaStream_t st) { GM_ROUTE_N(exact, L, st) }

// NUTS kernels are instantiated for EPL 4, 8, 25 (f32) / 13 (f64) with runtime masking; 2-D targets use EPL 2.
bool choose_nuts_decomposition(int dim, int dtype, int kind, int* epl, int* lpc) {
  if (dim <= 0 || kind == 1) return false;
  if (kind == 2 || kind == 4) {
    if (dim != 2) return false;
    *epl = 2; *lpc = 1;
    return true;
  }
  const int big = dtype == 0 ? 25 : 13;
  const int menu[3] = {4, 8, big};
  {
    const char* e_env = std::getenv("GMCMC_NUTS_EPL");
    const char* l_env = std::getenv("GMCMC_NUTS_LPC");
    if (e_env && l_env) {
      const int e = std::atoi(e_env), l = std::atoi(l_env);
      if ((e == 4 || e == 8 || e == big) && l >= 1 && l <= 32 && (l & (l - 1)) == 0 && e * l >= dim) {
        *epl = e; *lpc = l;
        return true;
      }
    }
  }
  // NUTS iterations are latency-bound (dependent global / shuffle / transcendental chains), so chains are spread over
  // MANY lanes with few coordinates each (short instruction streams, low register count, more chains in flight):
  // 8 coordinates per lane while that fits a warp, the wide slices beyond.  Measured on B200 at d = 100: (8, 16)
  // 4.6e8 leapfrogs/s, (25, 4) 4.0e8, (4, 32) 2.7e8.
  if (dim <= 4) { *epl = 4; *lpc = 1; return true; }
  for (int l = 1; l <= 32; l <<= 1)
    if (8 * l >= dim) { *epl = 8; *lpc = l; return true; }
  for (int l = 16; l <= 32; l <<= 1)
    if (big * l >= dim) { *epl = big; *lpc = l; return true; }
  return false;
}

// Decomposition: minimise padded slots (compute), penalising non exact fits (masking costs ~30 %) and
// lanes per chain; ties go to more elements per lane (more ILP, fewer shuffles).  Lanes past the end of
// the chain own no coordinate and only take part in the shuffles.  f64 is capped at 16 elements per lane
// (4 live arrays x 16 x 2 registers).  GMCMC_EPL / GMCMC_LPC override for tuning.
bool choose_decomposition(int dim, int dtype, int kind, int* epl, int* lpc) {
  if (dim <= 0) return false;
  if (kind == 1 || kind == 2 || kind == 4) {  // fixed 2-D targets: one lane per chain
    if (dim != 2) return false;
    *epl = 2; *lpc = 1;
    return true;
  }
  static const int menu[] = {1, 2, 3, 4, 8, 13, 16, 25, 32};
  const char* e_env = std::getenv("GMCMC_EPL");
  const char* l_env = std::getenv("GMCMC_LPC");
  if (e_env && l_env) {
    int e = std::atoi(e_env), l = std::atoi(l_env);
    bool ok = false;
    for (int m : menu) ok = ok || (m == e);
    if (ok && l >= 1 && l <= 32 && (l & (l - 1)) == 0 && e * l >= dim && !(dtype == 1 && e > 16)) {
      *epl = e; *lpc = l;
      return true;
    }
  }
  double best = 1e30;
  int be = 0, bl = 0;
  for (int l = 1; l <= 32; l <<= 1)
    for (int e : menu) {
      if (dtype == 1 && e > 16) continue;
      if (e * l < dim) continue;
      // padded slots cost compute (masking ~30 %); every extra lane costs shuffles in the stencil halo
      // and the reduction trees
      double cost = (double)e * l * ((e * l == dim) ? 1.0 : 1.3) + 4.0 * l;
      if (cost < best - 1e-9 || (cost < best + 1e-9 && e > be)) { best = cost; be = e; bl = l; }
    }
  if (!be) return false;
  *epl = be; *lpc = bl;
  return true;
}

}  // namespace gm
