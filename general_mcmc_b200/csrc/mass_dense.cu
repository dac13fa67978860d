// mass_dense.cu — dense mass-matrix update of the NUTS warm-up (sm_100a), compiled with --fmad=false.
//
// Restates, one thread per chain and in the reference's operation order,
//   maybe_update_mass_matrix, Dense branch        /root/reference/src/generic_nuts.rs:970-997
//   MassMatrix::dense_from_cov                    :208-226  (up to 8 tries, diagonal jitter x10 per try)
//   cholesky_spd / invert_spd_from_cholesky       :306-359
//   RunningCov::reset                             :99-107
// The covariance estimate (1 - reg) m2 / (n - 1) (+ reg on the diagonal, floored at the jitter) is formed on the fly
// from the running sums; the factor L and L^-1 live in chain-fastest work arrays [d, d, C] so the chains of a warp
// touch consecutive addresses.  Runs a handful of times per warm-up (window ends), never inside the sampling loop.
#include "kernels.h"

#include <cmath>

namespace gm {

namespace {

template <class T>
__global__ void __launch_bounds__(128) dense_mass_update_kernel(size_t C, int d, unsigned int n, T* __restrict__ run_mean,
                                                                T* __restrict__ run_m2, T* __restrict__ run_m2d,
                                                                T* __restrict__ inv_out, T* __restrict__ chol_out,
                                                                T* __restrict__ wl, T* __restrict__ wi, T reg, T jitter0,
                                                                int* __restrict__ state) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const T n_denom = (T)(n - 1u);
  const T omr = T(1) - reg;
  T* m2 = run_m2d + c * (size_t)d * d;
  auto L = [&](int i, int j) -> T& { return wl[((size_t)i * d + j) * C + c]; };
  auto IL = [&](int i, int j) -> T& { return wi[((size_t)i * d + j) * C + c]; };
  // covariance entry (i, j) as maybe_update_mass_matrix builds it (upper triangle of the running sums, mirrored)
  auto cov = [&](int i, int j) -> T {
    const int a = i < j ? i : j, b = i < j ? j : i;
    const T raw = m2[(size_t)a * d + b] / n_denom;
    if (a == b) { const T v = omr * raw + reg; return v > jitter0 ? v : jitter0; }
    return omr * raw;
  };
  T jit = jitter0 > (T)1e-10 ? jitter0 : (T)1e-10;
  bool ok = false;
  for (int attempt = 0; attempt < 8 && !ok; ++attempt) {
    ok = true;
    for (int i = 0; i < d && ok; ++i) {
      for (int j = 0; j <= i; ++j) {
        T sum = cov(i, j);
        if (i == j) sum = sum + jit;
        for (int k = 0; k < j; ++k) sum = sum - L(i, k) * L(j, k);
        if (i == j) {
          if (!(sum > T(0)) || !(fabs(sum) < T(INFINITY))) { ok = false; break; }
          L(i, j) = sqrt(sum);
        } else {
          const T dd = L(j, j);
          if (!(dd > T(0)) || !(fabs(dd) < T(INFINITY))) { ok = false; break; }
          L(i, j) = sum / dd;
        }
      }
    }
    if (!ok) jit = jit * (T)10.0;
  }
  T* inv = inv_out + c * (size_t)d * d;
  T* chol = chol_out + c * (size_t)d * d;
  if (ok) {
    // invert_spd_from_cholesky: L^-1 column by column, then inv = L^-T L^-1
    for (int i = 0; i < d; ++i) {
      IL(i, i) = T(1) / L(i, i);
      for (int j = i + 1; j < d; ++j) {
        T sum = T(0);
        for (int k = i; k < j; ++k) sum = sum + L(j, k) * IL(k, i);
        IL(j, i) = -sum / L(j, j);
      }
    }
    for (int i = 0; i < d; ++i) {
      for (int j = 0; j <= i; ++j) {
        T sum = T(0);
        for (int k = i; k < d; ++k) sum = sum + IL(k, i) * IL(k, j);      // k from max(i, j) = i
        inv[(size_t)i * d + j] = sum;
        inv[(size_t)j * d + i] = sum;
      }
    }
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) chol[(size_t)i * d + j] = j <= i ? L(i, j) : T(0);
    state[c] = 2;
  } else if (state[c] == 0) {
    // factorisation failed on every try and the chain still has the identity: diagonal_from_var(ones) (:989-994)
    for (int i = 0; i < d; ++i)
      for (int j = 0; j < d; ++j) { inv[(size_t)i * d + j] = i == j ? T(1) : T(0); chol[(size_t)i * d + j] = i == j ? T(1) : T(0); }
    state[c] = 1;
  }   // else: keep the current mass matrix
  // RunningCov::reset
  for (int i = 0; i < d; ++i) {
    run_mean[c * d + i] = T(0);
    run_m2[c * d + i] = T(0);
    for (int j = 0; j < d; ++j) m2[(size_t)i * d + j] = T(0);
  }
}

}  // namespace

cudaError_t launch_dense_mass_update(const DenseMassUpdate& U, cudaStream_t st) {
  const unsigned blocks = (unsigned)((U.n_chains + 127) / 128);
  if (U.dtype == 0)
    dense_mass_update_kernel<float><<<blocks, 128, 0, st>>>(U.n_chains, U.d, U.n, (float*)U.run_mean, (float*)U.run_m2, (float*)U.run_m2d,
                                                            (float*)U.inv, (float*)U.chol, (float*)U.scratch_l, (float*)U.scratch_invl,
                                                            (float)U.regularize, (float)U.jitter, U.state);
  else
    dense_mass_update_kernel<double><<<blocks, 128, 0, st>>>(U.n_chains, U.d, U.n, (double*)U.run_mean, (double*)U.run_m2, (double*)U.run_m2d,
                                                             (double*)U.inv, (double*)U.chol, (double*)U.scratch_l, (double*)U.scratch_invl,
                                                             U.regularize, U.jitter, U.state);
  return cudaGetLastError();
}

}  // namespace gm
