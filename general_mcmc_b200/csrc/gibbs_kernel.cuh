// gibbs_kernel.cuh — Gibbs sampling sweeps for many chains (sm_100a).  Compile with --fmad=false.
//
// Replaces the host loop of GibbsMarkovChain::step (/root/reference/src/gibbs.rs:89-105: for each coordinate i in
// 0 .. d, state[i] = conditional.sample(i, &state)) run per chain by ChainRunner::run (core.rs:95-115, 219-229; f64
// samples [chains, samples, dim]).  The reference's `Conditional<S>` (distributions.rs `Conditional::sample(&mut self, i,
// given) -> S`) is a host closure with its own RNG; on the device it is a struct with a static `sample` function,
// either one of the built-ins below (the conditionals of the reference's own tests, gibbs.rs:177-245) or a user-written
// one compiled ahead of time into a plugin (gmcmc_custom_conditional.cuh):
//
//   struct MyConditional {
//     static constexpr int dim = 2;
//     template <class RNG>
//     __device__ static double sample(int i, const double (&given)[dim], const double* params, RNG& rng);
//   };
//
// `rng.normal()` / `rng.uniform()` return N(0,1) / U[0,1) draws.  RNG contract: draw k of coordinate i at transition s of
// global chain g comes from Philox block (g, s, stream 0, block 4 i + k), k < 4: its words (0, 1) make the normal
// (Box-Muller, cosine branch, 53-bit-free: u1 = (w0 + 1) 2^-32, u2 = w1 2^-32), its words (2, 3) the uniform
// ((w2 2^32 + w3) >> 11) 2^-53; the normal and uniform counters of a coordinate advance separately.  With injected
// streams (gmcmc_gibbs_inject) the FIRST normal and the FIRST uniform of every (transition, coordinate) are read from the
// injected arrays instead — that is what the per-step parity tests against the oracle use.
// One thread per chain, the state in registers / local memory, f64 arithmetic in the reference's operation order.
#pragma once
#include "kernels.h"
#include "philox.cuh"

#include <cmath>

namespace gm {

struct GibbsRng {
  PhiloxKey key;
  unsigned long long gchain;
  uint32_t step;
  int coord, kn, ku;
  const double* inj_n;   // this (transition, chain)'s injected normals [d], or null
  const double* inj_u;

  __device__ __forceinline__ void begin(int i) { coord = i; kn = ku = 0; }
  __device__ __forceinline__ double normal() {
    const int k = kn++;
    if (inj_n && k == 0) return inj_n[coord];
    const uint4 r = philox4x32_10(philox_ctr(gchain, step, 0u, (uint32_t)(4 * coord + (k & 3))), key);
    const double u1 = ((double)r.x + 1.0) * 2.3283064365386963e-10;   // (0, 1]
    const double u2 = (double)r.y * 2.3283064365386963e-10;           // [0, 1)
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
  __device__ __forceinline__ double uniform() {
    const int k = ku++;
    if (inj_u && k == 0) return inj_u[coord];
    const uint4 r = philox4x32_10(philox_ctr(gchain, step, 0u, (uint32_t)(4 * coord + (k & 3))), key);
    const unsigned long long w = ((unsigned long long)r.z << 32) | r.w;
    return (double)(w >> 11) * 1.1102230246251565e-16;                // [0, 1)
  }
};

// ---- the conditionals of the reference's tests (gibbs.rs:177-245)
// ConstantConditional { c }: every coordinate becomes c (any dim <= kGibbsMaxDim)
template <int DIM>
struct CondConstant {
  static constexpr int dim = DIM;
  template <class RNG>
  __device__ static double sample(int, const double (&)[DIM], const double* params, RNG&) { return params[0]; }
};

// MixtureConditional { mu0, sigma0, mu1, sigma1, pi0 } on the state [x, z], z in {0.0, 1.0}:
//   x | z ~ N(mu_z, sigma_z^2);   P(z = 1 | x) = p1 / (p0 + p1), p_k = pi_k normal_pdf(x; mu_k, sigma_k)   (gibbs.rs:200-243)
struct CondMixtureXZ {
  static constexpr int dim = 2;
  __device__ static double normal_pdf(double x, double mu, double sigma) {
    const double var = sigma * sigma;
    const double coeff = 1.0 / sqrt(2.0 * 3.14159265358979323846 * var);
    const double dx = x - mu;
    const double exp_val = exp(-(dx * dx) / (2.0 * var));
    return coeff * exp_val;
  }
  template <class RNG>
  __device__ static double sample(int i, const double (&given)[2], const double* p, RNG& rng) {
    const double mu0 = p[0], sigma0 = p[1], mu1 = p[2], sigma1 = p[3], pi0 = p[4];
    if (i == 0) {
      const double z = given[1];
      const double noise = rng.normal();
      return z < 0.5 ? mu0 + sigma0 * noise : mu1 + sigma1 * noise;
    }
    const double x = given[0];
    const double p0 = pi0 * normal_pdf(x, mu0, sigma0);
    const double p1 = (1.0 - pi0) * normal_pdf(x, mu1, sigma1);
    const double total = p0 + p1;
    const double prob_z1 = total > 0.0 ? p1 / total : 0.5;
    return rng.uniform() < prob_z1 ? 1.0 : 0.0;
  }
};

template <class COND>
__global__ void __launch_bounds__(128) gibbs_run_kernel(const GibbsLaunch a) {
  constexpr int D = COND::dim;
  const size_t chain = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= a.n_chains) return;
  double x[D];
#pragma unroll
  for (int i = 0; i < D; ++i) x[i] = a.state[chain * D + i];
  GibbsRng rng;
  rng.key = PhiloxKey{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
  rng.gchain = a.chain_offset + chain;
  for (uint32_t s = 0; s < a.n_steps; ++s) {
    rng.step = a.step_base + s;
    rng.inj_n = a.inj_normals ? a.inj_normals + ((size_t)s * a.n_chains + chain) * D : nullptr;
    rng.inj_u = a.inj_uniforms ? a.inj_uniforms + ((size_t)s * a.n_chains + chain) * D : nullptr;
    // one full sweep (gibbs.rs:96-99): coordinate i sees the already-updated coordinates 0 .. i - 1
#pragma unroll
    for (int i = 0; i < D; ++i) {
      rng.begin(i);
      x[i] = COND::sample(i, x, a.params, rng);
    }
    if (a.out && s >= a.n_skip) {
      double* o = a.out + (chain * a.out_n + a.out_t0 + (s - a.n_skip)) * (size_t)D;
#pragma unroll
      for (int i = 0; i < D; ++i) o[i] = x[i];
    }
  }
#pragma unroll
  for (int i = 0; i < D; ++i) a.state[chain * D + i] = x[i];
}

template <class COND>
inline cudaError_t gibbs_launch(const GibbsLaunch& L, cudaStream_t st) {
  if (L.dim != COND::dim) return cudaErrorInvalidValue;
  const unsigned blocks = (unsigned)((L.n_chains + 127) / 128);
  gibbs_run_kernel<COND><<<blocks, 128, 0, st>>>(L);
  return cudaGetLastError();
}

}  // namespace gm
