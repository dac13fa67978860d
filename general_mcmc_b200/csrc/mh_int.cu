// mh_int.cu — integer-state Metropolis–Hastings (sm_100a), compiled with --fmad=false.
//
// Replaces the host loop of MetropolisHastings<S = i32, T = f64> (/root/reference/src/metropolis_hastings.rs:306-318,
// core.rs:95-115) for the discrete targets and +-1 random-walk proposals of the reference's own discrete-state tests
// (/root/reference/tests/metrohast_poisson_test.rs:18-86 Poisson, :195-252 Binomial):
//   unnorm_logp(k) = k ln(lambda) - lambda - ln(k!)                     (Poisson;  -inf for k < 0)
//                  = ln C(n, k) + k ln p + (n - k) ln(1 - p)            (Binomial; -inf outside [0, n])
//   proposal: k +- 1 with probability 1/2 each, clamped to the support (the reference reflects k = -1 to 0 and clamps
//   the binomial walk to [0, n]); Proposal::logp == ln 0.5 in both directions, kept in the ratio as the reference does.
// One thread per chain, state in registers, f64 arithmetic in the reference's operation order.  ln(k!) comes from a table
// built on the host with the reference's own summation loop (bit-identical to the CPU value) and is continued on the device
// beyond the table.  RNG contract: stream 0, block 0: bit i of the 128-bit block = direction of coordinate i (1: +1);
// stream 1, block 0: accept uniform from (r0, r1) as in the float-state kernel.
#include "kernels.h"
#include "philox.cuh"

#include <cmath>

namespace gm {

namespace {

constexpr int kMaxIntDim = 8;

struct IntTargetDev {
  int kind, dim, n;
  double lambda, ln_lambda, ln_p, ln_1mp, ln_half;
  const double* lnfact;   // [n_tab]: ln(k!) summed as the reference does
  int n_tab;
};

__device__ __forceinline__ double ln_factorial_dev(const IntTargetDev& t, int k) {
  if (k < 2) return 0.0;
  if (k < t.n_tab) return t.lnfact[k];
  double acc = t.lnfact[t.n_tab - 1];
  for (int i = t.n_tab; i <= k; ++i) acc += log((double)i);
  return acc;
}

__device__ __forceinline__ double int_logp(const IntTargetDev& t, const int (&k)[kMaxIntDim]) {
  double total = 0.0;
  for (int i = 0; i < t.dim; ++i) {
    double lp;
    if (t.kind == 0) {
      if (k[i] < 0) return -INFINITY;
      const double kf = (double)k[i];
      lp = kf * t.ln_lambda - t.lambda - ln_factorial_dev(t, k[i]);
    } else {
      if (k[i] < 0 || k[i] > t.n) return -INFINITY;
      const double kf = (double)k[i], nf = (double)t.n;
      const double coeff = ln_factorial_dev(t, t.n) - ln_factorial_dev(t, k[i]) - ln_factorial_dev(t, t.n - k[i]);
      lp = coeff + kf * t.ln_p + (nf - kf) * t.ln_1mp;
    }
    total = (i == 0) ? lp : total + lp;
  }
  return total;
}

__global__ void __launch_bounds__(128) mh_int_run_kernel(const IntTargetDev t, const MhIntLaunch a) {
  const size_t chain = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= a.n_chains) return;
  const unsigned long long gchain = a.chain_offset + chain;
  const PhiloxKey key{(uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
  const int d = t.dim;
  int x[kMaxIntDim];
  for (int i = 0; i < kMaxIntDim; ++i) x[i] = i < d ? a.state[chain * d + i] : 0;
  unsigned int n_accept = 0;
  for (uint32_t s = 0; s < a.n_steps; ++s) {
    const uint32_t step = a.step_base + s;
    int prop[kMaxIntDim];
    uint4 bits = make_uint4(0u, 0u, 0u, 0u);
    if (!a.inj_steps) bits = philox4x32_10(philox_ctr(gchain, step, 0u, 0u), key);
    for (int i = 0; i < kMaxIntDim; ++i) {
      prop[i] = 0;
      if (i < d) {
        int dir;
        if (a.inj_steps) dir = (int)a.inj_steps[((size_t)s * a.n_chains + chain) * d + i];
        else dir = ((bits.x >> i) & 1u) ? 1 : -1;
        const int v = x[i] + dir;
        prop[i] = t.kind == 0 ? (v < 0 ? 0 : v) : min(max(v, 0), t.n);
      }
    }
    const double lp_cur = int_logp(t, x);
    const double lp_prop = int_logp(t, prop);
    const double log_ratio = (lp_prop + t.ln_half) - (lp_cur + t.ln_half);
    double ln_u;
    if (a.inj_lnu) ln_u = a.inj_lnu[(size_t)s * a.n_chains + chain];
    else ln_u = log(accept_uniform<double>(philox4x32_10(philox_ctr(gchain, step, 1u, 0u), key)));
    const bool accept = log_ratio > ln_u;
    if (accept) { for (int i = 0; i < kMaxIntDim; ++i) x[i] = prop[i]; ++n_accept; }
    if (a.diag_logratio) { a.diag_logratio[(size_t)s * a.n_chains + chain] = log_ratio; a.diag_acc[(size_t)s * a.n_chains + chain] = accept ? 1 : 0; }
    if (a.out && s >= a.n_skip) {
      double* o = a.out + (chain * a.out_n + a.out_t0 + (s - a.n_skip)) * (size_t)d;
      for (int i = 0; i < d; ++i) o[i] = (double)x[i];
    }
  }
  for (int i = 0; i < d; ++i) a.state[chain * d + i] = x[i];
  for (int o = 16; o > 0; o >>= 1) n_accept += __shfl_xor_sync(__activemask(), n_accept, o);
  if ((threadIdx.x & 31) == 0 && n_accept) atomicAdd(a.accept_total, (unsigned long long)n_accept);
}

}  // namespace

int mh_int_max_dim() { return kMaxIntDim; }

cudaError_t launch_mh_int(const MhIntLaunch& L, cudaStream_t st) {
  IntTargetDev t;
  t.kind = L.kind; t.dim = L.dim; t.n = L.n;
  t.lambda = L.lambda; t.ln_lambda = L.ln_lambda; t.ln_p = L.ln_p; t.ln_1mp = L.ln_1mp; t.ln_half = L.ln_half;
  t.lnfact = L.lnfact; t.n_tab = L.n_tab;
  const unsigned blocks = (unsigned)((L.n_chains + 127) / 128);
  mh_int_run_kernel<<<blocks, 128, 0, st>>>(t, L);
  return cudaGetLastError();
}

}  // namespace gm
