// gmcmc_custom_conditional.cuh — ahead-of-time compiled conditionals for the Gibbs sampler (plugin interface).
//
// The reference's GibbsSampler takes any `Conditional<S>` (/root/reference/src/distributions.rs `Conditional::sample`,
// used by gibbs.rs:89-105).  On the device a user conditional is a struct with a static `sample` function (contract in
// gibbs_kernel.cuh), compiled with nvcc into a small shared library that instantiates the sweep kernel for it:
//
//   #include "gmcmc_custom_conditional.cuh"
//   struct BivariateNormal {            // x_i | x_j ~ N(rho x_j, 1 - rho^2)
//     static constexpr int dim = 2;
//     template <class RNG>
//     __device__ static double sample(int i, const double (&given)[dim], const double* params, RNG& rng) {
//       const double rho = params[0];
//       return rho * given[1 - i] + sqrt(1.0 - rho * rho) * rng.normal();
//     }
//   };
//   GMCMC_REGISTER_CONDITIONAL(BivariateNormal)
//
//   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -Xcompiler -fPIC -shared \
//        -I <repo>/general_mcmc_b200/csrc -I <repo>/include cond.cu -o libcond.so
//
// and is handed to gmcmc_gibbs_create_custom(ctx, "libcond.so", params, n_params, ...) (general_mcmc_b200.CustomConditional).
#pragma once
#include "gibbs_kernel.cuh"

#define GMCMC_REGISTER_CONDITIONAL(U)                                                                        \
  static_assert(U::dim >= 1 && U::dim <= 64, "conditional dim must be 1..64");                                \
  extern "C" const gm::CustomConditionalVTable* gmcmc_conditional_entry(void) {                               \
    static const gm::CustomConditionalVTable vt = {gm::kCustomAbiVersion, U::dim, &gm::gibbs_launch<U>};       \
    return &vt;                                                                                               \
  }
