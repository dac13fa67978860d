// gmcmc_custom_target.cuh — ahead-of-time compiled custom targets (plugin interface).
//
// The reference lets users implement `GradientTarget` / `BatchedGradientTarget`
// (/root/reference/src/distributions.rs:67-90) as host closures differentiated by burn autodiff.  On the
// device a custom target is a user-written `__device__` log-density-and-gradient function, compiled ahead of
// time with nvcc into a small shared library that instantiates the same fused kernels (K1 trajectory kernel,
// gradient evaluation, K5 NUTS, K2 Metropolis-Hastings) for it and exposes them through one C entry point:
//
//   #include "gmcmc_custom_target.cuh"
//   struct Banana {
//     static constexpr int dim = 2;
//     // log density at x (returned) and its gradient (written to g); params = the doubles handed to
//     // gmcmc_target_create_custom, converted to T, in device memory
//     template <class T>
//     __device__ static T logp_grad(const T (&x)[dim], T (&g)[dim], const T* params) { ... }
//   };
//   GMCMC_REGISTER_CUSTOM_TARGET(Banana)
//
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -Xcompiler -fPIC -shared \
//        -I <repo>/general_mcmc_b200/csrc -I <repo>/include banana.cu -o libbanana.so
//
// and is registered with gmcmc_target_create_custom(ctx, "libbanana.so", dtype, params, n, &target)
// (general_mcmc_b200.CustomTarget in Python).  One thread per chain, dim <= 32.
#pragma once
#include "nuts_kernel.cuh"   // pulls hmc_kernel.cuh
#define GM_MH_NO_BUILTIN_LAUNCHERS
#include "mh_kernel.cuh"     // K2 kernel template (MetropolisHastings with a custom Target)

namespace gm {
namespace GM_NS {

template <class U>
struct TagCustom {};

template <class T, int EPL, bool PADDED, bool WANT_LOGP, class U>
__device__ __forceinline__ T eval_target(TagCustom<U>, const T (&x)[EPL], T (&g)[EPL], const Lane&, const TParams<T>& tp, T*) {
  static_assert(EPL == U::dim, "custom targets run with one thread per chain");
  return U::template logp_grad<T>(x, g, tp.dp);
}

template <class U>
inline cudaError_t custom_launch_hmc(const HmcLaunch& L, cudaStream_t st) {
  if (L.epl != U::dim || L.lpc != 1 || L.tgt.dim != U::dim) return cudaErrorInvalidValue;
  if (L.tgt.dtype == 0) return launch_one<float, U::dim, TagCustom<U>, false>(L, st);
  return launch_one<double, U::dim, TagCustom<U>, false>(L, st);
}
template <class U>
inline cudaError_t custom_launch_eval(const EvalLaunch& E, cudaStream_t st) {
  if (E.epl != U::dim || E.lpc != 1 || E.tgt.dim != U::dim) return cudaErrorInvalidValue;
  if (E.tgt.dtype == 0) return eval_one<float, U::dim, TagCustom<U>, false>(E, st);
  return eval_one<double, U::dim, TagCustom<U>, false>(E, st);
}
template <class U>
inline cudaError_t custom_launch_nuts(const NutsLaunch& L, cudaStream_t st) {
  if (L.epl != U::dim || L.lpc != 1 || L.tgt.dim != U::dim) return cudaErrorInvalidValue;
  if (L.tgt.dtype == 0) return nuts_launch_one<float, U::dim, TagCustom<U>>(L, st);
  return nuts_launch_one<double, U::dim, TagCustom<U>>(L, st);
}

// MetropolisHastings on a custom target (Target::unnorm_logp, distributions.rs:107-110): the K2 kernel with the
// plugin's log density (the gradient its function also writes is dead code here and is removed by the compiler).
template <class U>
struct CustomLogp {
  template <class T>
  __device__ static __forceinline__ T logp(const T (&x)[U::dim], const T* params) {
    T g[U::dim];
    return U::template logp_grad<T>(x, g, params);
  }
};
template <class T, class U>
inline cudaError_t custom_mh_one(const MhLaunch& L, cudaStream_t st) {
  MhArgs<T> a = make_mh_args<T>(L);
  const unsigned blocks = (unsigned)((L.n_chains + kMhBlock - 1) / kMhBlock);
  if (a.inj_normals || a.inj_lnu || a.diag_logratio)
    mh_run_kernel<T, U::dim, kTargetCustom, true, true, CustomLogp<U>><<<blocks, kMhBlock, 0, st>>>(a);
  else
    mh_run_kernel<T, U::dim, kTargetCustom, true, false, CustomLogp<U>><<<blocks, kMhBlock, 0, st>>>(a);
  return cudaGetLastError();
}
template <class U>
inline cudaError_t custom_launch_mh(const MhLaunch& L, cudaStream_t st) {
  if (L.tgt.dim != U::dim) return cudaErrorInvalidValue;
  return L.tgt.dtype == 0 ? custom_mh_one<float, U>(L, st) : custom_mh_one<double, U>(L, st);
}

}  // namespace GM_NS
}  // namespace gm

#define GMCMC_REGISTER_CUSTOM_TARGET(U)                                                                   \
  static_assert(U::dim >= 1 && U::dim <= 32, "custom target dim must be 1..32");                            \
  extern "C" const gm::CustomTargetVTable* gmcmc_custom_entry(void) {                                      \
    static const gm::CustomTargetVTable vt = {gm::kCustomAbiVersion, U::dim, &gm::GM_NS::custom_launch_hmc<U>,      \
                                              &gm::GM_NS::custom_launch_eval<U>, &gm::GM_NS::custom_launch_nuts<U>,  \
                                              &gm::GM_NS::custom_launch_mh<U>};                                       \
    return &vt;                                                                                             \
  }
