// hmc_kernel.cuh — K1: fused HMC trajectory kernel (sm_100a).
//
// Replaces, as ONE kernel per run segment, the ~40 allocating tensor passes per leapfrog of
//   BatchedGenericHMC::step / leapfrog      /root/reference/src/batched_hmc.rs:129-190
//   GenericHMC::step / leapfrog_chain       /root/reference/src/generic_hmc.rs:166-221
//   BatchVector ops for Tensor<B,2>         /root/reference/src/euclidean.rs:447-534
//   autodiff logp_and_grad adapter          /root/reference/src/hmc.rs:42-61
//   HMC::run stack + permute                /root/reference/src/hmc.rs:164-181
//
// Mapping: a chain is spread over `lpc` adjacent lanes of one warp (lpc = 1,2,..,32); each lane owns
// EPL contiguous coordinates in registers.  q, p and the gradient never leave registers during the L
// leapfrog steps; cross-lane traffic is two shuffles per gradient (stencil halo) and one shuffle tree
// per reduction (kinetic energy, log density).  A per-warp shared-memory row per chain stages (a) the
// Philox normals (generated block-aligned by the chain's lanes, contract in include/gmcmc.h) and
// (b) the accepted position, so the [chains, samples, dim] write-out is coalesced and vectorised.
//
// This header is compiled twice: GM_EXACT=0 (default flags: FMA contraction, merged half kicks,
// log-density only where the Hamiltonian needs it) and GM_EXACT=1 with nvcc --fmad=false (reference
// operation order, sequential left-to-right sums: bit-for-bit the CPU oracle).
#pragma once
#include "kernels.h"
#include "philox.cuh"

#include <cmath>

#ifndef GM_EXACT
#define GM_EXACT 0
#endif
#ifndef GM_MINB
#define GM_MINB 3   // resident CTAs per SM the trajectory kernel is register-budgeted for (tuned on B200)
#endif
#if GM_EXACT
#define GM_NS exact
#else
#define GM_NS fast
#endif

namespace gm {
namespace GM_NS {

constexpr bool kExact = (GM_EXACT != 0);
constexpr unsigned kFull = 0xffffffffu;

template <class T> struct VecOf;
template <> struct VecOf<float> { using type = float4; static constexpr int n = 4; };
template <> struct VecOf<double> { using type = double2; static constexpr int n = 2; };


struct Lane {
  int part;       // which slice of the chain this lane owns
  int lpc;        // lanes per chain
  int gbase;      // warp lane index of part 0 of this chain
  int lo;         // first coordinate owned
  int d;          // chain dimension
  int nvalid;     // number of owned coordinates that are < d
  bool first;     // owns coordinate 0
  bool last;      // owns coordinate d-1 as its LAST slot (exact-fit decompositions only)
};

template <int EPL>
__device__ __forceinline__ Lane make_lane(int tid_in_grid, int lpc, int d) {
  Lane ln;
  ln.lpc = lpc;
  ln.part = tid_in_grid & (lpc - 1);
  ln.gbase = (threadIdx.x & 31) - ln.part;
  ln.lo = ln.part * EPL;
  ln.d = d;
  int nv = d - ln.lo;
  ln.nvalid = nv < 0 ? 0 : (nv > EPL ? EPL : nv);
  ln.first = ln.part == 0;
  ln.last = (ln.lo + EPL == d);
  return ln;
}

// Sum of per-lane term arrays over the whole chain; every lane of the chain gets the result.  Slots
// past the end of the chain must hold exact zeros (adding +0 never changes a sum), so there is no mask.
// EXACT: strict left-to-right order over coordinates (lane k continues lane k-1's running sum).
template <class T, int EPL>
__device__ __forceinline__ T chain_sum(const T (&terms)[EPL], const Lane& ln) {
  if constexpr (kExact) {
    T s = T(0);
    for (int k = 0; k < ln.lpc; ++k) {
      if (ln.part == k) {
#pragma unroll
        for (int j = 0; j < EPL; ++j) s = s + terms[j];
      }
      s = __shfl_sync(kFull, s, ln.gbase + k);
    }
    return s;
  } else {
    T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      if ((j & 3) == 0) s0 += terms[j];
      else if ((j & 3) == 1) s1 += terms[j];
      else if ((j & 3) == 2) s2 += terms[j];
      else s3 += terms[j];
    }
    T s = (s0 + s1) + (s2 + s3);
    for (int o = ln.lpc >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    return s;
  }
}

template <class T>
__device__ __forceinline__ T chain_max(T v, const Lane& ln) {
  for (int o = ln.lpc >> 1; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// ----------------------------------------------------------------------------------------------
// Targets.  eval<WANT_LOGP>(x, g, ...) writes the gradient slice and returns the chain's log
// density (valid on every lane of the chain) when WANT_LOGP, else 0.
// ----------------------------------------------------------------------------------------------
template <class T>
struct TParams {          // per-launch target parameters in T
  T sp[kMaxScalarParams];
  const T* dp;            // device block (dense Gaussian: mu[d], P[d*d], nc; mixture: w[K], mu[K*d], ln w[K])
  const T* smem_mu;       // mixture means staged in shared memory, lane-padded [K][roundup(EPL, 4) / VN][lpc][VN] (16-byte units), or null
  const T* smem_logw;     // mixture log weights staged in shared memory [K], or null
  int n_comp;
  // derived for DiffableGaussian2D (distributions.rs:229-253)
  T inv_cov[2][2];
  T norm_const;
};

struct TagRosenbrockND {};
struct TagIsoGauss {};
struct TagDenseGauss {};
struct TagMixture {};
struct TagRosenbrock2D {};
struct TagDiffGauss2D {};
struct TagGauss2D {};

// RosenbrockND — distributions.rs:544-554:  logp = -sum_{i<d-1} [100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2]
// Gradient (the reference differentiates the same expression by autodiff):
//   g_i = [i<d-1] (400 t_i x_i + 2 (1 - x_i)) + [i>0] (-200 t_{i-1}),  t_i = x_{i+1} - x_i^2
template <class T, int EPL, bool PADDED, bool WANT_LOGP>
__device__ __forceinline__ T eval_target(TagRosenbrockND, const T (&x)[EPL], T (&g)[EPL], const Lane& ln,
                                         const TParams<T>&, T*) {
  const T x_next_lane = __shfl_down_sync(kFull, x[0], 1);
  T gt[EPL];     // fast: t_j ; exact: gt_j = -200 t_j
  T terms[EPL];
  bool low[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const T xn = (j + 1 < EPL) ? x[j + 1] : x_next_lane;
    bool has_low;
    if constexpr (PADDED) has_low = (ln.lo + j) < (ln.d - 1);
    else has_low = (j + 1 < EPL) ? true : !ln.last;
    low[j] = has_low;
    T t = xn - x[j] * x[j];
    if (!has_low) t = T(0);
    if constexpr (kExact) {
      T u = (-x[j]) + T(1);
      terms[j] = has_low ? ((t * t) * T(100) + u * u) : T(0);
      gt[j] = T(-200) * t;
    } else {
      if constexpr (WANT_LOGP) {
        T u = T(1) - x[j];
        terms[j] = has_low ? (T(100) * t * t + u * u) : T(0);
      }
      gt[j] = t;
    }
  }
  T prev = __shfl_up_sync(kFull, gt[EPL - 1], 1);
  if (ln.first) prev = T(0);
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const T gp = (j == 0) ? prev : gt[j - 1];
    if constexpr (kExact) {
      T u = (-x[j]) + T(1);
      T gi;
      if (low[j]) {
        gi = (T(-2) * (gt[j] * x[j])) + T(2) * u;
        if (ln.lo + j > 0) gi = gi + gp;
      } else {
        gi = (ln.lo + j > 0 && ln.lo + j < ln.d) ? gp : T(0);
      }
      g[j] = gi;
    } else {
      T a = T(400) * gt[j] - T(2);
      T lowterm = x[j] * a + T(2);
      if (!low[j]) lowterm = T(0);
      T gi = T(-200) * gp + lowterm;
      if constexpr (PADDED) { if (ln.lo + j >= ln.d) gi = T(0); }
      g[j] = gi;
    }
  }
  if constexpr (WANT_LOGP || kExact) {
    if constexpr (!WANT_LOGP) return T(0);
    // number of coordinates with a "low" term in this lane: i < d-1
    int nl = ln.d - 1 - ln.lo;
    nl = nl < 0 ? 0 : (nl > EPL ? EPL : nl);
    return -chain_sum<T, EPL>(terms, ln);
  }
  return T(0);
}

// IsotropicGaussian as a gradient target — distributions.rs:398-406 (nuts.rs:484-496 StandardNormal)
template <class T, int EPL, bool PADDED, bool WANT_LOGP>
__device__ __forceinline__ T eval_target(TagIsoGauss, const T (&x)[EPL], T (&g)[EPL], const Lane& ln,
                                         const TParams<T>& tp, T*) {
  const T var = tp.sp[0] * tp.sp[0];
  T terms[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    terms[j] = (j < ln.nvalid) ? (x[j] * x[j]) : T(0);
    g[j] = (j < ln.nvalid) ? (-x[j] / var) : T(0);
  }
  if constexpr (!WANT_LOGP) return T(0);
  T sum = chain_sum<T, EPL>(terms, ln);
  return -T(0.5) * sum / var;
}

// Dense-covariance Gaussian, generic-d register/shared path (the tensor-core path for large d is K3):
//   delta = x - mu ; z = delta . P ; logp = nc - 0.5 sum(z * delta) ; grad = -z   (P symmetric)
// N-D form of DiffableGaussian2D::unnorm_logp_batch, distributions.rs:265-291.
template <class T, int EPL, bool PADDED, bool WANT_LOGP>
__device__ __forceinline__ T eval_target(TagDenseGauss, const T (&x)[EPL], T (&g)[EPL], const Lane& ln,
                                         const TParams<T>& tp, T* row) {
  const int d = ln.d;
  const T* mu = tp.dp;
  const T* P = tp.dp + d;
  T delta[EPL];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    delta[j] = (j < ln.nvalid) ? (x[j] - mu[ln.lo + j]) : T(0);
    if (j < ln.nvalid) row[ln.lo + j] = delta[j];
  }
  __syncwarp();
  T z[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) z[j] = T(0);
  for (int i = 0; i < d; ++i) {
    const T di = row[i];
    const T* Pi = P + (size_t)i * d + ln.lo;
#pragma unroll
    for (int j = 0; j < EPL; ++j)
      if (j < ln.nvalid) z[j] = z[j] + di * Pi[j];
  }
  __syncwarp();
  T terms[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    g[j] = -z[j];
    terms[j] = z[j] * delta[j];
  }
  if constexpr (!WANT_LOGP) return T(0);
  T quad = chain_sum<T, EPL>(terms, ln);
  const T nc = tp.dp[(size_t)d + (size_t)d * d];
  return nc - quad * T(0.5);
}

// Isotropic Gaussian mixture (synthetic target of BASELINE cfg5):
//   logp = logsumexp_k( ln w_k - |x - mu_k|^2 / (2 sigma^2) ),  grad = sum_k r_k (mu_k - x) / sigma^2
constexpr int kMaxComp = 8;
// The component loops are deliberately NOT unrolled (K is a runtime value): unrolling 8 components x EPL
// coordinates made the NUTS kernel ~200 KB of code and instruction-fetch bound.
// fast-math transcendental helpers: MUFU approximations in fast mode (f32), libdevice otherwise
template <class T> __device__ __forceinline__ T fast_log(T x) {
  if constexpr (!kExact && sizeof(T) == 4) return __logf(x); else return log(x);
}
template <class T> __device__ __forceinline__ T fast_exp(T x) {
  if constexpr (!kExact && sizeof(T) == 4) return __expf(x); else return exp(x);
}
template <class T> __device__ __forceinline__ T fast_div(T a, T b) {
  if constexpr (!kExact && sizeof(T) == 4) return __fdividef(a, b); else return a / b;
}

// this lane's slice of component k's mean: from the lane-padded shared-memory copy (vector loads) or from HBM / L1
template <class T, int EPL>
__device__ __forceinline__ void mixture_mean_slice(T (&m)[EPL], const TParams<T>& tp, const T* mu, int k, int d, const Lane& ln) {
  if (tp.smem_mu) {
    using V = typename VecOf<T>::type;
    constexpr int VN = VecOf<T>::n;
    constexpr int EPLP = (EPL + 3) / 4 * 4;
    const V* s = reinterpret_cast<const V*>(tp.smem_mu + (size_t)k * ln.lpc * EPLP) + ln.part;   // unit i of lane `part` at i * lpc + part
#pragma unroll
    for (int i = 0; i < EPLP / VN; ++i) {
      if (i * VN < EPL) {
        const V v = s[i * ln.lpc];
        if constexpr (VN == 4) {
          m[i * 4] = v.x;
          if (i * 4 + 1 < EPL) m[i * 4 + 1 < EPL ? i * 4 + 1 : 0] = v.y;
          if (i * 4 + 2 < EPL) m[i * 4 + 2 < EPL ? i * 4 + 2 : 0] = v.z;
          if (i * 4 + 3 < EPL) m[i * 4 + 3 < EPL ? i * 4 + 3 : 0] = v.w;
        } else {
          m[i * 2] = v.x;
          if (i * 2 + 1 < EPL) m[i * 2 + 1 < EPL ? i * 2 + 1 : 0] = v.y;
        }
      }
    }
  } else {
    const T* muk = mu + (size_t)k * d + ln.lo;
#pragma unroll
    for (int j = 0; j < EPL; ++j) m[j] = (j < ln.nvalid) ? muk[j] : T(0);
  }
}

template <class T, int EPL, bool PADDED, bool WANT_LOGP>
__device__ __forceinline__ T eval_target(TagMixture, const T (&x)[EPL], T (&g)[EPL], const Lane& ln,
                                         const TParams<T>& tp, T*) {
  const int K = tp.n_comp;
  const int d = ln.d;
  const T sigma = tp.sp[1];
  const T inv_var = T(1) / (sigma * sigma);
  const T* mu = tp.dp + K;
  const T* logw = tp.smem_logw ? tp.smem_logw : (tp.dp + K + (size_t)K * d);   // ln w[K], host-computed (runtime.cu)
  T a[kMaxComp];          // dynamically indexed: lives in (L1-resident) local memory
  T amax = -INFINITY;
  bool packed = false;
#ifndef GM_MIX_NO_PACKED_REDUCE
  if constexpr (!kExact) packed = (K == 4 && ln.lpc >= 4);   // fast math mode only: the summation order changes
#endif
  if (packed) {
    // Four components, chains spread over >= 4 lanes: the four per-lane partial sums are reduced together by a
    // transpose-reduce (2 + 1 shuffles leave one component per lane, log2(lpc) - 2 finish it, 1 + 2 gather the four
    // totals back) — 8 shuffles in 7 dependent steps at lpc = 16 instead of 4 x 4 in sequence.
    T sk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      T m[EPL];
      mixture_mean_slice<T, EPL>(m, tp, mu, k, d, ln);
      T acc0 = T(0), acc1 = T(0);
#pragma unroll
      for (int j = 0; j < EPL; ++j) {
        const T df = (!PADDED || j < ln.nvalid) ? (x[j] - m[j]) : T(0);
        if (j & 1) acc1 = fma(df, df, acc1); else acc0 = fma(df, df, acc0);
      }
      sk[k] = acc0 + acc1;
    }
    const int ln32 = (int)(threadIdx.x & 31u);
    const bool b0 = (ln32 & 1) != 0, b1 = (ln32 & 2) != 0;
    T keep0 = b0 ? sk[2] : sk[0], keep1 = b0 ? sk[3] : sk[1];
    keep0 += __shfl_xor_sync(kFull, b0 ? sk[0] : sk[2], 1);
    keep1 += __shfl_xor_sync(kFull, b0 ? sk[1] : sk[3], 1);
    T keep = b1 ? keep1 : keep0;
    keep += __shfl_xor_sync(kFull, b1 ? keep0 : keep1, 2);          // this lane now owns component 2 b0 + b1
    for (int o = 4; o < ln.lpc; o <<= 1) keep += __shfl_xor_sync(kFull, keep, o);
    const T other = __shfl_xor_sync(kFull, keep, 2);
    const T q0 = b1 ? other : keep, q1 = b1 ? keep : other;        // components 2 b0, 2 b0 + 1
    const T o0 = __shfl_xor_sync(kFull, q0, 1), o1 = __shfl_xor_sync(kFull, q1, 1);
    sk[0] = b0 ? o0 : q0; sk[1] = b0 ? o1 : q1; sk[2] = b0 ? q0 : o0; sk[3] = b0 ? q1 : o1;
    // K = 4 stays in registers and fully unrolled from here on (no local-memory array, no component loops)
    T ak[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ak[k] = logw[k] - T(0.5) * sk[k] * inv_var;
    }
    const T am = max(max(ak[0], ak[1]), max(ak[2], ak[3]));
    T ek[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) ek[k] = fast_exp<T>(ak[k] - am);
    const T se4 = (ek[0] + ek[1]) + (ek[2] + ek[3]);
    const T inv_se = fast_div<T>(T(1), se4);
    T acc4[EPL];
#pragma unroll
    for (int j = 0; j < EPL; ++j) acc4[j] = T(0);
    // grad = sum_k r_k (mu_k - x) / sigma^2 = (sum_k r_k mu_k - x) / sigma^2 (the responsibilities sum to one to within the
    // rounding of the approximate division): one FMA per component and coordinate instead of a subtraction and an FMA
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      T m[EPL];
      mixture_mean_slice<T, EPL>(m, tp, mu, k, d, ln);
      const T rk = ek[k] * inv_se;
#pragma unroll
      for (int j = 0; j < EPL; ++j) acc4[j] = fma(rk, m[j], acc4[j]);
    }
#pragma unroll
    for (int j = 0; j < EPL; ++j) g[j] = (!PADDED || j < ln.nvalid) ? (acc4[j] - x[j]) * inv_var : T(0);
    return am + fast_log<T>(se4);
  } else {
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      T m[EPL], terms[EPL];
      mixture_mean_slice<T, EPL>(m, tp, mu, k, d, ln);
#pragma unroll
      for (int j = 0; j < EPL; ++j) {
        const T df = (!PADDED || j < ln.nvalid) ? (x[j] - m[j]) : T(0);
        terms[j] = df * df;
      }
      const T sq = chain_sum<T, EPL>(terms, ln);
      const T ak = logw[k] - T(0.5) * sq * inv_var;
      a[k] = ak;
      amax = max(amax, ak);
    }
  }
  T se = T(0);
#pragma unroll 1
  for (int k = 0; k < K; ++k) { const T e = fast_exp<T>(a[k] - amax); a[k] = e; se = se + e; }
  T acc[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) acc[j] = T(0);
#pragma unroll 1
  for (int k = 0; k < K; ++k) {
    T m[EPL];
    mixture_mean_slice<T, EPL>(m, tp, mu, k, d, ln);
    const T rk = fast_div<T>(a[k], se);
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      const T dm = (!PADDED || j < ln.nvalid) ? (m[j] - x[j]) : T(0);
      acc[j] = acc[j] + rk * dm;
    }
  }
#pragma unroll
  for (int j = 0; j < EPL; ++j) g[j] = acc[j] * inv_var;
  return amax + fast_log<T>(se);
}

// Rosenbrock2D — distributions.rs:502-515 (one lane per chain, EPL == 2)
template <class T, int EPL, bool PADDED, bool WANT_LOGP>
__device__ __forceinline__ T eval_target(TagRosenbrock2D, const T (&x)[EPL], T (&g)[EPL], const Lane&,
                                         const TParams<T>& tp, T*) {
  static_assert(EPL == 2, "Rosenbrock2D is 2-D");
  const T a = tp.sp[0], b = tp.sp[1];
  T u = (-x[0]) + a;
  T t = x[1] - x[0] * x[0];
  T term1 = u * u;
  T term2 = (t * t) * b;
  T gt = -((T(2) * t) * b);
  g[1] = gt;
  g[0] = (T(2) * u) + (-(T(2) * (gt * x[0])));
  return -(term1 + term2);
}

// DiffableGaussian2D — distributions.rs:265-291 (gradient in backward-pass order, see oracle)
template <class T, int EPL, bool PADDED, bool WANT_LOGP>
__device__ __forceinline__ T eval_target(TagDiffGauss2D, const T (&x)[EPL], T (&g)[EPL], const Lane&,
                                         const TParams<T>& tp, T*) {
  static_assert(EPL == 2, "DiffableGaussian2D is 2-D");
  T d0 = x[0] - tp.sp[0], d1 = x[1] - tp.sp[1];
  T z0 = d0 * tp.inv_cov[0][0] + d1 * tp.inv_cov[1][0];
  T z1 = d0 * tp.inv_cov[0][1] + d1 * tp.inv_cov[1][1];
  T quad = z0 * d0 + z1 * d1;
  const T h = T(0.5);
  T m0 = -(h * d0), m1 = -(h * d1);
  g[0] = (-(h * z0)) + (m0 * tp.inv_cov[0][0] + m1 * tp.inv_cov[0][1]);
  g[1] = (-(h * z1)) + (m0 * tp.inv_cov[1][0] + m1 * tp.inv_cov[1][1]);
  return tp.norm_const - quad * h;
}

// Gaussian2D — distributions.rs:195-207 (+ gradient, which the reference does not define)
template <class T, int EPL, bool PADDED, bool WANT_LOGP>
__device__ __forceinline__ T eval_target(TagGauss2D, const T (&x)[EPL], T (&g)[EPL], const Lane&,
                                         const TParams<T>& tp, T*) {
  static_assert(EPL == 2, "Gaussian2D is 2-D");
  T a = tp.sp[2], b = tp.sp[3], c = tp.sp[4], dd = tp.sp[5];
  T det = a * dd - b * c;
  T d0 = x[0] - tp.sp[0], d1 = x[1] - tp.sp[1];
  T i00 = dd / det, i01 = (-b) / det, i10 = (-c) / det, i11 = a / det;
  T v0 = d0 * i00 + d1 * i10;
  T v1 = d0 * i01 + d1 * i11;
  T w0 = i00 * d0 + i01 * d1;
  T w1 = i10 * d0 + i11 * d1;
  g[0] = -T(0.5) * (v0 + w0);
  g[1] = -T(0.5) * (v1 + w1);
  return -T(0.5) * (v0 * d0 + v1 * d1);
}

// ----------------------------------------------------------------------------------------------
// eval_kick: gradient at x, then the momentum kick p += coef * g applied `NK` times (NK = 2 is the
// reference's two consecutive half-kicks with the same gradient, batched_hmc.rs:187 + :175).  Returns
// the log density when WANT_LOGP.  Generic version: gradient slice in registers, then the kick.
// ----------------------------------------------------------------------------------------------
// Gradient cache (fast mode, RosenbrockND): a transition of L leapfrogs needs L + 1 gradients, and the first of them is at
// the current point — the gradient the previous transition already evaluated at its trajectory end (accepted) or start
// (rejected).  The kernel keeps both in two shared-memory rows per chain and swaps them on acceptance, so a transition
// evaluates L gradients; the cached value is the same function of the same position, hence the same bits.
// An experiment, OFF by default: measured on B200 (65,536 chains, d = 100, L = 32, all parity tests green) it is 0.6 % SLOWER
// (51.97 vs 51.68 us per transition, profiles/r2_k1_gradcache_experiment.txt) although it issues 2.5 % fewer instructions:
// 168 instead of 162 registers and 25 LDS + 25 STS per transition on the trajectory's critical path cost what the
// 125 saved FFMAs bought.
#ifndef GM_K1_GRADCACHE
#define GM_K1_GRADCACHE 0
#endif
template <class TAG> struct GradCache { static constexpr bool on = false; };
template <> struct GradCache<TagRosenbrockND> { static constexpr bool on = GM_K1_GRADCACHE && !kExact; };

template <class T, int EPL, bool PADDED, bool WANT_LOGP, int NK, class TAG>
__device__ __forceinline__ T eval_kick(TAG, const T (&x)[EPL], T (&p)[EPL], const T coef, const Lane& ln,
                                       const TParams<T>& tp, T* row) {
  T g[EPL];
  const T lp = eval_target<T, EPL, PADDED, WANT_LOGP>(TAG{}, x, g, ln, tp, row);
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    p[j] = p[j] + g[j] * coef;
    if constexpr (NK == 2) p[j] = p[j] + g[j] * coef;
  }
  return lp;
}

// RosenbrockND: the gradient never materialises — each coordinate's stencil value goes straight into
// its momentum (6 FMAs per coordinate per leapfrog in fast mode: drift 1, t 1, gradient 3, kick 1).
template <class T, int EPL, bool PADDED, bool WANT_LOGP, int NK, bool SAVE = false>
__device__ __forceinline__ T eval_kick(TagRosenbrockND, const T (&x)[EPL], T (&p)[EPL], const T coef, const Lane& ln,
                                       const TParams<T>&, T*, T* gsave = nullptr /* SAVE: the lane's slice of a gradient row */) {
  const T x_next_lane = __shfl_down_sync(kFull, x[0], 1);
  T gt[EPL];     // fast: t_j ; exact: gt_j = -200 t_j
  T terms[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const T xn = (j + 1 < EPL) ? x[j + 1] : x_next_lane;
    bool has_low;
    if constexpr (PADDED) has_low = (ln.lo + j) < (ln.d - 1);
    else has_low = (j + 1 < EPL) ? true : !ln.last;
    T t = xn - x[j] * x[j];
    if (!has_low) t = T(0);
    if constexpr (kExact) {
      if constexpr (WANT_LOGP) {
        T u = (-x[j]) + T(1);
        terms[j] = has_low ? ((t * t) * T(100) + u * u) : T(0);
      }
      gt[j] = T(-200) * t;
    } else {
      if constexpr (WANT_LOGP) {
        T u = T(1) - x[j];
        terms[j] = has_low ? (T(100) * t * t + u * u) : T(0);
      }
      gt[j] = t;
    }
  }
  T prev = __shfl_up_sync(kFull, gt[EPL - 1], 1);
  if (ln.first) prev = T(0);
#pragma unroll
  for (int j = 0; j < EPL; ++j) {
    const T gp = (j == 0) ? prev : gt[j - 1];
    bool has_low;
    if constexpr (PADDED) has_low = (ln.lo + j) < (ln.d - 1);
    else has_low = (j + 1 < EPL) ? true : !ln.last;
    T gi;
    if constexpr (kExact) {
      T u = (-x[j]) + T(1);
      if (has_low) {
        gi = (T(-2) * (gt[j] * x[j])) + T(2) * u;
        if (ln.lo + j > 0) gi = gi + gp;
      } else {
        gi = (ln.lo + j > 0 && ln.lo + j < ln.d) ? gp : T(0);
      }
    } else {
      T a = T(400) * gt[j] - T(2);
      T lowterm = x[j] * a + T(2);
      if (!has_low) lowterm = T(0);
      gi = T(-200) * gp + lowterm;
      if constexpr (PADDED) { if (ln.lo + j >= ln.d) gi = T(0); }
    }
    p[j] = p[j] + gi * coef;
    if constexpr (NK == 2) p[j] = p[j] + gi * coef;
    if constexpr (SAVE) { if (!PADDED || j < ln.nvalid) gsave[j] = gi; }
  }
  if constexpr (WANT_LOGP) {
    int nl = ln.d - 1 - ln.lo;
    nl = nl < 0 ? 0 : (nl > EPL ? EPL : nl);
    return -chain_sum<T, EPL>(terms, ln);
  }
  return T(0);
}

// ------------------------------------------------------------------------------------------------
// RosenbrockND, f32, fast mode, exact-fit decomposition: the (L - 1) inner leapfrogs on PACKED pairs — an experiment,
// OFF by default because it measured slower.
// The scalar loop above is bound by the instruction issue rate (FFMA is 77 % of the issued instructions, issue slots
// 84 % busy, FMA pipe 73 %: profiles/r2_hmc_run_kernel_full.txt).  Blackwell's fma.rn.f32x2 (SASS FFMA2) does two FMAs
// per issued instruction, so the same six FMAs per coordinate need half the issue slots: 108 instead of 160 instructions
// per leapfrog (72 FFMA2 with immediate / broadcast operands, 22 MOV, 6 FFMA).  Coordinates j and j + H (H = EPL / 2) of a
// lane form pair j: the "next coordinate" of a pair is then simply the next pair, and the stencil needs no re-packing
// except at the lane's two ends.  The loop tracks y = -q: with s_j = y_j^2 + y_{j+1} (= -t_j), a' = 400 s + 2,
// low = y a' + 2, g = 200 s_{j-1} + low every operation is a plain a * b + c (f32x2 has no operand negation), and each
// result is the exact negative / equal of the scalar formulation's (round-to-nearest is symmetric), so the trajectory is
// bit-identical to the scalar loop's (the whole GPU suite passes with it on).
// Measured on B200 (65,536 chains, d = 100, L = 32): 57.0 us per transition against 53.4 us for the scalar loop
// (3.62e10 vs 3.85e10 grad-evals/s): the packed form reads the same registers per FMA, and register-file reads are
// the limit (tools/microbench_fp32.cu: FFMA2 with three distinct register pairs sustains 47 TFLOP/s like scalar
// three-register FFMA, against 72 for constant operands), so halving the instruction count buys nothing.
// ------------------------------------------------------------------------------------------------
// tuning knob: unroll factor of the inner leapfrog loop (default: the compiler's choice, 2)
#define GM_PRAGMA_(x) _Pragma(#x)
#define GM_PRAGMA(x) GM_PRAGMA_(x)
#ifdef GM_K1_LOOP_UNROLL
#define GM_K1_UNROLL_PRAGMA GM_PRAGMA(unroll GM_K1_LOOP_UNROLL)
#else
#define GM_K1_UNROLL_PRAGMA
#endif
#ifndef GM_K1_PACKED
#define GM_K1_PACKED 0
#endif
#ifndef GM_K1_LOAD_UN
#define GM_K1_LOAD_UN 8    // loads in flight per lane of the launch-head state load (padded rows; 8 and 32 measure the same)
#endif
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ float lo2(unsigned long long v) { float lo, hi; upk2(v, lo, hi); return lo; }
__device__ __forceinline__ float hi2(unsigned long long v) { float lo, hi; upk2(v, lo, hi); return hi; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// n_iter times: q += eps p ; p += eps grad(q)   (the merged-kick inner loop of the fast mode)
template <int EPL>
__device__ __forceinline__ void rosen_leapfrogs_packed(float (&q)[EPL], float (&p)[EPL], const float eps, const uint32_t n_iter,
                                                       const Lane& ln) {
  constexpr int H = EPL / 2;
  constexpr bool ODD = (EPL & 1) != 0;
  unsigned long long y2[H], p2[H];
  float ys = 0.f, ps = 0.f;          // the odd coordinate 2H
#pragma unroll
  for (int k = 0; k < H; ++k) { y2[k] = pk2(-q[k], -q[k + H]); p2[k] = pk2(p[k], p[k + H]); }
  if constexpr (ODD) { ys = -q[2 * H]; ps = p[2 * H]; }
  const float neps = -eps;
  const unsigned long long c_neps = pk2(neps, neps), c_eps = pk2(eps, eps);
  const unsigned long long c400 = pk2(400.f, 400.f), c200 = pk2(200.f, 200.f), c2 = pk2(2.f, 2.f);
  for (uint32_t it = 0; it < n_iter; ++it) {
    // drift: y -= eps p
#pragma unroll
    for (int k = 0; k < H; ++k) y2[k] = fma2(p2[k], c_neps, y2[k]);
    if constexpr (ODD) ys = fmaf(ps, neps, ys);
    const float y_next_lane = __shfl_down_sync(kFull, lo2(y2[0]), 1);
    // s_j = y_j^2 + y_{j+1}; the last coordinate of the chain has no such term
    unsigned long long s2[H];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k + 1 < H; ++k) s2[k] = fma2(y2[k], y2[k], y2[k + 1]);
    if constexpr (ODD) {
      s2[H - 1] = fma2(y2[H - 1], y2[H - 1], pk2(hi2(y2[0]), ys));
      ss = fmaf(ys, ys, y_next_lane);
      if (ln.last) ss = 0.f;
    } else {
      s2[H - 1] = fma2(y2[H - 1], y2[H - 1], pk2(hi2(y2[0]), y_next_lane));
      if (ln.last) s2[H - 1] = pk2(lo2(s2[H - 1]), 0.f);
    }
    float prev = __shfl_up_sync(kFull, ODD ? ss : hi2(s2[H - 1]), 1);
    if (ln.first) prev = 0.f;
    // gradient straight into the momentum: a' = 400 s + 2, low = y a' + 2, g = 200 s_{j-1} + low, p += eps g
#pragma unroll
    for (int k = 0; k < H; ++k) {
      const unsigned long long a2 = fma2(s2[k], c400, c2);
      unsigned long long low = fma2(y2[k], a2, c2);
      if constexpr (!ODD) { if (k == H - 1 && ln.last) low = pk2(lo2(low), 0.f); }
      const unsigned long long sp = (k == 0) ? pk2(prev, lo2(s2[H - 1])) : s2[k - 1];
      const unsigned long long g2 = fma2(sp, c200, low);
      p2[k] = fma2(g2, c_eps, p2[k]);
    }
    if constexpr (ODD) {
      const float a1 = fmaf(ss, 400.f, 2.f);
      float low = fmaf(ys, a1, 2.f);
      if (ln.last) low = 0.f;
      const float g1 = fmaf(hi2(s2[H - 1]), 200.f, low);
      ps = fmaf(g1, eps, ps);
    }
  }
#pragma unroll
  for (int k = 0; k < H; ++k) {
    float a, b;
    upk2(y2[k], a, b); q[k] = -a; q[k + H] = -b;
    upk2(p2[k], a, b); p[k] = a; p[k + H] = b;
  }
  if constexpr (ODD) { q[2 * H] = -ys; p[2 * H] = ps; }
}

template <class T>
__host__ inline TParams<T> make_tparams(const TargetDesc& td) {
  TParams<T> tp;
  for (int i = 0; i < kMaxScalarParams; ++i) tp.sp[i] = (T)td.sp[i];
  tp.dp = (const T*)td.dparams;
  tp.smem_mu = nullptr;
  tp.smem_logw = nullptr;
  tp.n_comp = td.n_comp;
  tp.inv_cov[0][0] = tp.inv_cov[0][1] = tp.inv_cov[1][0] = tp.inv_cov[1][1] = T(0);
  tp.norm_const = T(0);
  if (td.kind == 2 /* DIFF_GAUSS2D */) {
    // DiffableGaussian2D::new, distributions.rs:229-253, evaluated in T on the host (IEEE, same as oracle)
    volatile T c00 = tp.sp[2], c01 = tp.sp[3], c10 = tp.sp[4], c11 = tp.sp[5];
    volatile T m1 = c00 * c11;
    volatile T m2 = c01 * c10;
    volatile T det_cov = m1 - m2;
    volatile T inv_det = T(1) / det_cov;
    tp.inv_cov[0][0] = c11 * inv_det;
    tp.inv_cov[0][1] = -c01 * inv_det;
    tp.inv_cov[1][0] = -c10 * inv_det;
    tp.inv_cov[1][1] = c00 * inv_det;
    T logdet = std::log((T)det_cov);
    T two = T(1) + T(1);
    const T pi = (T)3.14159265358979323846264338327950288;
    volatile T l2pi = std::log(two * pi);
    volatile T tl = two * l2pi;
    volatile T sum = tl + logdet;
    tp.norm_const = -sum / two;
  }
  return tp;
}

// ----------------------------------------------------------------------------------------------
// The trajectory kernel
// ----------------------------------------------------------------------------------------------
template <class T>
struct HmcArgs {
  TParams<T> tp;
  size_t n_chains;
  unsigned long long chain_offset;
  PhiloxKey key;
  uint32_t step_base;
  T* positions;
  const T* eps;         // device step size (null: use eps_val)
  T eps_val;            // step size known on the host at launch time (constant-bank operand)
  int eps_stride;
  int d, d_pad, lpc;
  uint32_t L, n_steps, n_skip;
  T* out;
  size_t out_n;
  uint32_t out_t0;
  unsigned long long* accept_total;
  unsigned long long* diverge_total;
  double* alpha_part;   // [n_steps][warps of the grid] per-transition, per-warp sums of min(1, exp(log_accept)), or null
  T* da_eps; T* da_eps_bar; T* da_h_bar; T* da_mu;
  uint32_t da_m_base, da_n_adapt;
  T da_delta;
  const T* inj_normals;
  const T* inj_lnu;
  T* diag_logacc;
  uint8_t* diag_acc;
  T* diag_pq;
  T* diag_pp;
};


// Write-out plan of one lane: the warp's position rows ([chains_in_warp][d_pad] in shared memory) are
// copied to [chain, slot, :] of the sample tensor as a flat list of vectors (float4 / double2, or
// scalars when d is not a multiple of the vector width); lane l owns list items l, l+32, ...  The
// shared-memory index and the global offset (relative to the warp's first chain, slot 0) of each item
// do not depend on the transition, so they are computed once per launch.
template <class T, int EPL>
struct StorePlan {
  static constexpr int VN = VecOf<T>::n;
  static constexpr int KMAX = (EPL + VN - 1) / VN + 1;   // vector items per lane (upper bound)
  static constexpr int KMAX_S = EPL + 1;                  // scalar items per lane (upper bound)
  bool vec;
  int n_items;            // total items of the warp
  unsigned soff[KMAX];    // shared-memory element index
  unsigned goff[KMAX];    // global element offset
};

template <class T, int EPL>
__device__ __forceinline__ void make_store_plan(StorePlan<T, EPL>& pl, int d, int d_pad, int chains_in_warp,
                                                size_t first_chain, size_t n_chains, size_t out_n) {
  constexpr int VN = StorePlan<T, EPL>::VN;
  const int lane = threadIdx.x & 31;
  size_t live = n_chains > first_chain ? n_chains - first_chain : 0;
  const int nch = live < (size_t)chains_in_warp ? (int)live : chains_in_warp;
  pl.vec = (d % VN) == 0 && (out_n * (size_t)d * chains_in_warp) < 0xffffffffull;
  if (pl.vec) {
    const int per = d / VN;
    pl.n_items = nch * per;
#pragma unroll
    for (int k = 0; k < StorePlan<T, EPL>::KMAX; ++k) {
      const int f = lane + 32 * k;
      const int c = f / per, i = f - c * per;
      pl.soff[k] = (unsigned)(c * d_pad + i * VN);
      pl.goff[k] = (unsigned)((size_t)c * out_n * d + (size_t)i * VN);
    }
  } else {
    pl.n_items = nch * d;
  }
}

template <class T, int EPL>
__device__ __forceinline__ void store_rows(const StorePlan<T, EPL>& pl, const T* warp_rows, int d, int d_pad,
                                           T* out_warp /* &out[first_chain, slot, 0] */, size_t out_n) {
  const int lane = threadIdx.x & 31;
  using V = typename VecOf<T>::type;
  if (pl.vec) {
    V v[StorePlan<T, EPL>::KMAX];
#pragma unroll
    for (int k = 0; k < StorePlan<T, EPL>::KMAX; ++k)
      if (lane + 32 * k < pl.n_items) v[k] = *reinterpret_cast<const V*>(warp_rows + pl.soff[k]);
#pragma unroll
    for (int k = 0; k < StorePlan<T, EPL>::KMAX; ++k)
      if (lane + 32 * k < pl.n_items) __stcs(reinterpret_cast<V*>(out_warp + pl.goff[k]), v[k]);
  } else {
    for (int f = lane; f < pl.n_items; f += 32) {
      const int c = f / d, i = f - c * d;
      __stcs(out_warp + (size_t)c * out_n * d + i, warp_rows[c * d_pad + i]);
    }
  }
}

// Shared memory per warp: [chains_in_warp][d_pad] current (accepted) positions — the row that is
// streamed to the sample tensor — followed by the same shape of scratch (Philox normals / dense
// target staging).  Registers hold only the moving point q and the momentum p.
template <class T, int EPL, class TAG, bool PADDED>
__global__ void __launch_bounds__(kHmcBlock, GM_MINB) hmc_run_kernel(const HmcArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);

  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const Lane ln = make_lane<EPL>(tid, a.lpc, a.d);
  const size_t chain = (size_t)(tid / a.lpc);
  const bool active = chain < a.n_chains;
  const int lane = threadIdx.x & 31;
  const int chains_in_warp = 32 / a.lpc;
  const int warp_in_block = threadIdx.x >> 5;
  const int chain_in_warp = lane / a.lpc;
  const size_t warp_elems = (size_t)chains_in_warp * a.d_pad;
  constexpr bool kGC = GradCache<TAG>::on;
  T* warp_pos = smem + (size_t)warp_in_block * (kGC ? 4 : 2) * warp_elems;
  T* warp_scr = warp_pos + warp_elems;
  // gradient cache: two more rows per chain, the lane's slice of row `g_sel` = gradient at the current point
  T* const g_rows = warp_scr + warp_elems + (size_t)chain_in_warp * a.d_pad + ln.lo;
  int g_sel = 0;
  T* pos_row = warp_pos + (size_t)chain_in_warp * a.d_pad;
  T* row = warp_scr + (size_t)chain_in_warp * a.d_pad;
  const size_t warp_first_chain = (size_t)((tid & ~31) / a.lpc);
  const unsigned long long gchain = a.chain_offset + chain;

  // current positions -> shared rows.  The warp's chains are one contiguous run of the [C, d] array: it is read as a flat
  // list with all loads in flight per lane (a dependent load -> store loop over the rows cost ~15 us of HBM latency per CTA,
  // more than a whole transition, which is what made short launches 2.3x as expensive per transition); rows of chains past
  // the end and padding columns get a harmless point.
  {
    const size_t live_chains = a.n_chains > warp_first_chain ? a.n_chains - warp_first_chain : 0;
    const int nch = live_chains < (size_t)chains_in_warp ? (int)live_chains : chains_in_warp;
    const int total = nch * a.d;
    const T* src = a.positions + warp_first_chain * (size_t)a.d;
    if (a.d_pad != a.d || nch < chains_in_warp)
      for (int i = lane; i < (int)warp_elems; i += 32) warp_pos[i] = T(1);
    __syncwarp();
    using V = typename VecOf<T>::type;
    constexpr int VN = VecOf<T>::n;
    if (a.d_pad == a.d && (a.d % VN) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      // unpadded rows (config 4: d = 100): the shared rows ARE the flat list, moved in 16-byte units.  (The general path
      // below divides every index by d: those divisions were 11 % of a one-transition launch's instructions,
      // profiles/r2_hmc_single_transition_launch_after.txt.)
      const V* src_v = reinterpret_cast<const V*>(src);
      V* dst_v = reinterpret_cast<V*>(warp_pos);
      const int nv = total / VN;
      constexpr int UV = 8;
      for (int base = 0; base < nv; base += 32 * UV) {
        V v[UV];
#pragma unroll
        for (int u = 0; u < UV; ++u) {
          const int idx = base + lane + 32 * u;
          if (idx < nv) v[u] = __ldg(src_v + idx);
        }
#pragma unroll
        for (int u = 0; u < UV; ++u) {
          const int idx = base + lane + 32 * u;
          if (idx < nv) dst_v[idx] = v[u];
        }
      }
    } else {
      constexpr int UN = GM_K1_LOAD_UN;
      for (int base = 0; base < total; base += 32 * UN) {
        T v[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int idx = base + lane + 32 * u;
          v[u] = idx < total ? __ldg(src + idx) : T(1);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int idx = base + lane + 32 * u;
          if (idx < total) {
            const int c = idx / a.d;
            warp_pos[(size_t)c * a.d_pad + (idx - c * a.d)] = v[u];
          }
        }
      }
    }
  }
  __syncwarp();

  StorePlan<T, EPL> plan;
  make_store_plan<T, EPL>(plan, a.d, a.d_pad, chains_in_warp, warp_first_chain, a.n_chains, a.out_n);

  T eps = a.eps ? (active ? a.eps[a.eps_stride ? chain : 0] : T(0.01)) : a.eps_val;
  // per-chain dual averaging state (GMCMC_ADAPT_PER_CHAIN; generic_nuts.rs:882-924)
  T da_eps_bar = T(1), da_h_bar = T(0), da_mu = T(0);
  const bool per_chain_da = a.da_eps != nullptr;
  if (per_chain_da && active) {
    eps = a.da_eps[chain]; da_eps_bar = a.da_eps_bar[chain]; da_h_bar = a.da_h_bar[chain]; da_mu = a.da_mu[chain];
  }

  unsigned int n_accept = 0, n_diverge = 0;
  // log density of the current point, carried from transition to transition: it is the same function
  // of the same position the reference re-evaluates (batched_hmc.rs:138), hence bit-identical
  T logp_cur = T(0);
  bool have_logp_cur = false;

  for (uint32_t s = 0; s < a.n_steps; ++s) {
    const uint32_t step = a.step_base + s;
    T q[EPL], p[EPL];

    // ---- 1. momentum ~ N(0, I)   (batched_hmc.rs:131 / generic_hmc.rs:177)
    if (a.inj_normals) {
#pragma unroll
      for (int j = 0; j < EPL; ++j)
        p[j] = (active && j < ln.nvalid) ? a.inj_normals[((size_t)s * a.n_chains + chain) * a.d + ln.lo + j] : T(0);
    } else {
      constexpr int NPB = NormalsPerBlock<T>::value;
      const int nblocks = (a.d + NPB - 1) / NPB;
      for (int b = ln.part; b < nblocks; b += a.lpc) {
        T z[NPB];
        normals_from_block<kExact>(philox4x32_10(philox_ctr(gchain, step, 0u, (uint32_t)b), a.key), z);
        // d_pad is a multiple of 4, so a block is either entirely inside the row or entirely outside
        if (b * NPB < a.d_pad) {
          using V = typename VecOf<T>::type;
          V v;
          if constexpr (NPB == 4) v = V{z[0], z[1], z[2], z[3]};
          else v = V{z[0], z[1]};
          *reinterpret_cast<V*>(row + b * NPB) = v;
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < EPL; ++j) p[j] = (!PADDED || j < ln.nvalid) ? row[ln.lo + j] : T(0);
      __syncwarp();
    }

    // ---- 2. kinetic energy, then log density + first kick at the current point (batched_hmc.rs:134-138)
    T terms[EPL];
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      q[j] = (!PADDED || j < ln.nvalid) ? pos_row[ln.lo + j] : T(1);
      terms[j] = p[j] * p[j];
    }
    const T ke0 = chain_sum<T, EPL>(terms, ln) * T(0.5);

    // ---- 3. L leapfrog steps (batched_hmc.rs:166-190 / generic_hmc.rs:204-221)
    const T half = T(0.5) * eps;
    T logp0, logp1;
    if (a.L == 0) {
      T g[EPL];
      logp0 = logp1 = eval_target<T, EPL, PADDED, true>(TAG{}, q, g, ln, a.tp, row);
    } else if constexpr (kExact) {
      // reference order: (p += g half ; q += p eps ; g = grad(q) ; p += g half) x L, kicks never merged
      if (have_logp_cur) { eval_kick<T, EPL, PADDED, false, 1>(TAG{}, q, p, half, ln, a.tp, row); logp0 = logp_cur; }
      else logp0 = eval_kick<T, EPL, PADDED, true, 1>(TAG{}, q, p, half, ln, a.tp, row);
      for (uint32_t l = 0; l + 1 < a.L; ++l) {
#pragma unroll
        for (int j = 0; j < EPL; ++j) q[j] = q[j] + p[j] * eps;
        eval_kick<T, EPL, PADDED, false, 2>(TAG{}, q, p, half, ln, a.tp, row);
      }
#pragma unroll
      for (int j = 0; j < EPL; ++j) q[j] = q[j] + p[j] * eps;
      logp1 = eval_kick<T, EPL, PADDED, true, 1>(TAG{}, q, p, half, ln, a.tp, row);
    } else {
      // merged kicks: p += eps/2 g ; (q += eps p ; p += eps grad(q)) x (L-1) ; q += eps p ; p += eps/2 grad(q)
      if constexpr (kGC) {
        T* g_cur = g_rows + (g_sel ? warp_elems : 0);
        if (have_logp_cur) {
#pragma unroll
          for (int j = 0; j < EPL; ++j)
            if (!PADDED || j < ln.nvalid) p[j] = p[j] + g_cur[j] * half;
          logp0 = logp_cur;
        } else {
          logp0 = eval_kick<T, EPL, PADDED, true, 1, true>(TAG{}, q, p, half, ln, a.tp, row, g_cur);
        }
      } else {
        if (have_logp_cur) { eval_kick<T, EPL, PADDED, false, 1>(TAG{}, q, p, half, ln, a.tp, row); logp0 = logp_cur; }
        else logp0 = eval_kick<T, EPL, PADDED, true, 1>(TAG{}, q, p, half, ln, a.tp, row);
      }
      if constexpr (GM_K1_PACKED && std::is_same<TAG, TagRosenbrockND>::value && sizeof(T) == 4 && !PADDED && EPL >= 4) {
        rosen_leapfrogs_packed<EPL>(q, p, eps, a.L - 1, ln);
      } else {
        // The loop is bound by register-file reads (a three-register FFMA sustains 47 of the 72 TFLOP/s a constant-operand
        // FFMA does, tools/microbench_fp32.cu).  When the step size is the launch-wide value the host passed by value, the
        // drift and the kick take it as a constant-bank operand: 53.3 -> 51.0 us per transition at config 4's shape.
        auto inner = [&](const T eps_l) {
          GM_K1_UNROLL_PRAGMA
          for (uint32_t l = 0; l + 1 < a.L; ++l) {
#pragma unroll
            for (int j = 0; j < EPL; ++j) q[j] += eps_l * p[j];
            eval_kick<T, EPL, PADDED, false, 1>(TAG{}, q, p, eps_l, ln, a.tp, row);
          }
        };
        if (a.eps == nullptr && !per_chain_da) inner(a.eps_val); else inner(eps);
      }
#pragma unroll
      for (int j = 0; j < EPL; ++j) q[j] += eps * p[j];
      if constexpr (kGC) logp1 = eval_kick<T, EPL, PADDED, true, 1, true>(TAG{}, q, p, half, ln, a.tp, row, g_rows + (g_sel ? 0 : warp_elems));
      else logp1 = eval_kick<T, EPL, PADDED, true, 1>(TAG{}, q, p, half, ln, a.tp, row);
    }

    // ---- 4. Hamiltonian + Metropolis accept (batched_hmc.rs:148-162 / generic_hmc.rs:195-200)
#pragma unroll
    for (int j = 0; j < EPL; ++j) terms[j] = p[j] * p[j];
    const T ke1 = chain_sum<T, EPL>(terms, ln) * T(0.5);
    const T log_accept = (logp1 - logp0) + (ke0 - ke1);
    T ln_u;
    if (a.inj_lnu) {
      ln_u = active ? a.inj_lnu[(size_t)s * a.n_chains + chain] : T(0);
    } else {
      ln_u = log(accept_uniform<T>(philox4x32_10(philox_ctr(gchain, step, 1u, 0u), a.key)));
    }
    const bool accept = (ln_u <= log_accept);
    if (accept) {
#pragma unroll
      for (int j = 0; j < EPL; ++j)
        if (!PADDED || j < ln.nvalid) pos_row[ln.lo + j] = q[j];
    }
    logp_cur = accept ? logp1 : logp0;
    have_logp_cur = true;
    if constexpr (kGC) g_sel ^= accept ? 1 : 0;   // the trajectory end's gradient row becomes the current point's
    const bool finite = (log_accept == log_accept) && (fabs(log_accept) < T(INFINITY));
    const T alpha = finite ? min(T(1), exp(log_accept)) : (log_accept > T(0) ? T(1) : T(0));
    if (active && ln.part == 0) {
      n_accept += accept ? 1u : 0u;
      n_diverge += finite ? 0u : 1u;
    }
    if (a.alpha_part) {
      // acceptance statistic of THIS transition, summed over the warp's chains in a fixed order (xor tree: the same
      // value on every lane), one partial per (transition, warp): the pooled dual-averaging input (collective A1)
      double al = (active && ln.part == 0) ? (double)alpha : 0.0;
      for (int o = 16; o > 0; o >>= 1) al += __shfl_xor_sync(kFull, al, o);
      if (lane == 0) a.alpha_part[(size_t)s * ((size_t)gridDim.x * (kHmcBlock / 32)) + ((size_t)(blockIdx.x * blockDim.x + threadIdx.x) >> 5)] = al;
    }

    if (a.diag_logacc && active) {
      const size_t idx = (size_t)s * a.n_chains + chain;
      if (ln.part == 0) { a.diag_logacc[idx] = log_accept; a.diag_acc[idx] = accept ? 1 : 0; }
      if (a.diag_pq) {
#pragma unroll
        for (int j = 0; j < EPL; ++j)
          if (j < ln.nvalid) { a.diag_pq[idx * a.d + ln.lo + j] = q[j]; a.diag_pp[idx * a.d + ln.lo + j] = p[j]; }
      }
    }

    // ---- 5. per-chain dual averaging (optional; constants generic_nuts.rs:638-641)
    if (per_chain_da) {
      const uint32_t m = a.da_m_base + s + 1;
      if (m <= a.da_n_adapt) {
        const T mm = (T)m;
        T eta = T(1) / (T)(m + 10u);
        da_h_bar = (T(1) - eta) * da_h_bar + eta * (a.da_delta - alpha);
        eps = exp(da_mu - sqrt(mm) / T(0.05) * da_h_bar);
        eta = pow(mm, -T(0.75));
        da_eps_bar = exp((T(1) - eta) * log(da_eps_bar) + eta * log(eps));
        if (m == a.da_n_adapt) eps = da_eps_bar;
      }
    }

    // ---- 6. write-out [chain, slot, :] straight from the position rows (hmc.rs:173-180 stack+permute, fused)
    __syncwarp();
    if (s >= a.n_skip && a.out) {
      store_rows<T, EPL>(plan, warp_pos, a.d, a.d_pad,
                         a.out + (warp_first_chain * a.out_n + (size_t)a.out_t0 + (s - a.n_skip)) * (size_t)a.d, a.out_n);
      __syncwarp();
    }
  }

  // ---- state back to HBM (the same flat list)
  {
    const size_t live_chains = a.n_chains > warp_first_chain ? a.n_chains - warp_first_chain : 0;
    const int nch = live_chains < (size_t)chains_in_warp ? (int)live_chains : chains_in_warp;
    const int total = nch * a.d;
    T* dst = a.positions + warp_first_chain * (size_t)a.d;
    using V = typename VecOf<T>::type;
    constexpr int VN = VecOf<T>::n;
    if (a.d_pad == a.d && (a.d % VN) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      const V* src_v = reinterpret_cast<const V*>(warp_pos);
      V* dst_v = reinterpret_cast<V*>(dst);
      const int nv = total / VN;
#pragma unroll 4
      for (int idx = lane; idx < nv; idx += 32) dst_v[idx] = src_v[idx];
    } else {
#pragma unroll 4
      for (int idx = lane; idx < total; idx += 32) {
        const int c = idx / a.d;
        dst[idx] = warp_pos[(size_t)c * a.d_pad + (idx - c * a.d)];
      }
    }
  }
  if (active && per_chain_da && ln.part == 0) {
    a.da_eps[chain] = eps; a.da_eps_bar[chain] = da_eps_bar; a.da_h_bar[chain] = da_h_bar;
  }
  // counters: warp reduce, one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    n_accept += __shfl_xor_sync(kFull, n_accept, o);
    n_diverge += __shfl_xor_sync(kFull, n_diverge, o);
  }
  if (lane == 0) {
    if (n_accept) atomicAdd(a.accept_total, (unsigned long long)n_accept);
    if (n_diverge) atomicAdd(a.diverge_total, (unsigned long long)n_diverge);
  }
}

// logp / gradient of a batch of points through the same target code (gmcmc_target_logp_grad)
template <class T, int EPL, class TAG, bool PADDED>
__global__ void __launch_bounds__(kHmcBlock) eval_kernel(TParams<T> tp, size_t n, int d, int d_pad, int lpc,
                                                         const T* x, T* logp, T* grad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* stage = reinterpret_cast<T*>(smem_raw);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const Lane ln = make_lane<EPL>(tid, lpc, d);
  const size_t pt = (size_t)(tid / lpc);
  const bool active = pt < n;
  T* row = stage + (size_t)(threadIdx.x / lpc) * d_pad;
  T q[EPL], g[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) q[j] = (active && j < ln.nvalid) ? x[pt * d + ln.lo + j] : T(1);
  T lp = eval_target<T, EPL, PADDED, true>(TAG{}, q, g, ln, tp, row);
  if (active) {
    if (ln.part == 0) logp[pt] = lp;
    if (grad) {
#pragma unroll
      for (int j = 0; j < EPL; ++j)
        if (j < ln.nvalid) grad[pt * d + ln.lo + j] = g[j];
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Host-side dispatch
// ----------------------------------------------------------------------------------------------
template <class T>
inline HmcArgs<T> make_args(const HmcLaunch& L) {
  HmcArgs<T> a;
  a.tp = make_tparams<T>(L.tgt);
  a.n_chains = L.n_chains;
  a.chain_offset = L.chain_offset;
  a.key = PhiloxKey{(uint32_t)L.seed, (uint32_t)(L.seed >> 32)};
  a.step_base = L.step_base;
  a.positions = (T*)L.positions;
  a.eps = (const T*)L.eps;
  a.eps_val = (T)L.eps_val;
  a.eps_stride = L.eps_stride;
  a.d = L.tgt.dim;
  constexpr int VN = VecOf<T>::n;
  a.d_pad = ((L.tgt.dim + 3) / 4) * 4;
  (void)VN;
  a.lpc = L.lpc;
  a.L = L.n_leapfrog; a.n_steps = L.n_steps; a.n_skip = L.n_skip;
  a.out = (T*)L.out; a.out_n = L.out_n; a.out_t0 = L.out_t0;
  a.accept_total = L.accept_total; a.diverge_total = L.diverge_total; a.alpha_part = L.alpha_part;
  a.da_eps = (T*)L.da_eps; a.da_eps_bar = (T*)L.da_eps_bar; a.da_h_bar = (T*)L.da_h_bar; a.da_mu = (T*)L.da_mu;
  a.da_m_base = L.da_m_base; a.da_n_adapt = L.da_n_adapt; a.da_delta = (T)L.da_delta;
  a.inj_normals = (const T*)L.inj_normals; a.inj_lnu = (const T*)L.inj_lnu;
  a.diag_logacc = (T*)L.diag_logacc; a.diag_acc = L.diag_acc; a.diag_pq = (T*)L.diag_pq; a.diag_pp = (T*)L.diag_pp;
  return a;
}

template <class T, int EPL, class TAG, bool PADDED>
inline cudaError_t launch_one(const HmcLaunch& L, cudaStream_t st) {
  HmcArgs<T> a = make_args<T>(L);
  const size_t threads = L.n_chains * (size_t)L.lpc;
  const unsigned blocks = (unsigned)((threads + kHmcBlock - 1) / kHmcBlock);
  const size_t smem = (GradCache<TAG>::on ? 4 : 2) * (size_t)(kHmcBlock / L.lpc) * a.d_pad * sizeof(T);
  auto kern = hmc_run_kernel<T, EPL, TAG, PADDED>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<blocks, kHmcBlock, smem, st>>>(a);
  return cudaGetLastError();
}

template <class T, int EPL, class TAG, bool PADDED>
inline cudaError_t eval_one(const EvalLaunch& E, cudaStream_t st) {
  TParams<T> tp = make_tparams<T>(E.tgt);
  const int d = E.tgt.dim, d_pad = ((d + 3) / 4) * 4;
  const size_t threads = E.n * (size_t)E.lpc;
  const unsigned blocks = (unsigned)((threads + kHmcBlock - 1) / kHmcBlock);
  const size_t smem = (size_t)(kHmcBlock / E.lpc) * d_pad * sizeof(T);
  auto kern = eval_kernel<T, EPL, TAG, PADDED>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<blocks, kHmcBlock, smem, st>>>(tp, E.n, d, d_pad, E.lpc, (const T*)E.x, (T*)E.logp, (T*)E.grad);
  return cudaGetLastError();
}

// EPL menu of the register-resident kernels
#define GM_FOR_EPL(X) X(1) X(2) X(3) X(4) X(8) X(13) X(16) X(25) X(32)

template <class T, class TAG, bool ALLOW_EXACT_FIT>
inline cudaError_t dispatch_epl(const HmcLaunch& L, cudaStream_t st) {
  const bool padded = (L.epl * L.lpc != L.tgt.dim);
  switch (L.epl) {
#define GM_CASE(E)                                                                        \
  case E:                                                                                 \
    if constexpr (sizeof(T) == 8 && (E) > 16) return cudaErrorInvalidValue;               \
    else {                                                                                \
      if constexpr (ALLOW_EXACT_FIT) { if (!padded) return launch_one<T, E, TAG, false>(L, st); } \
      return launch_one<T, E, TAG, true>(L, st);                                          \
    }
    GM_FOR_EPL(GM_CASE)
#undef GM_CASE
  }
  return cudaErrorInvalidValue;
}

template <class T, class TAG, bool ALLOW_EXACT_FIT>
inline cudaError_t dispatch_eval_epl(const EvalLaunch& E, cudaStream_t st) {
  const bool padded = (E.epl * E.lpc != E.tgt.dim);
  switch (E.epl) {
#define GM_CASE(EP)                                                                      \
  case EP:                                                                                \
    if constexpr (sizeof(T) == 8 && (EP) > 16) return cudaErrorInvalidValue;              \
    else {                                                                                \
      if constexpr (ALLOW_EXACT_FIT) { if (!padded) return eval_one<T, EP, TAG, false>(E, st); } \
      return eval_one<T, EP, TAG, true>(E, st);                                           \
    }
    GM_FOR_EPL(GM_CASE)
#undef GM_CASE
  }
  return cudaErrorInvalidValue;
}

}  // namespace GM_NS
}  // namespace gm
