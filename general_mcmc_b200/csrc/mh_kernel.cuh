// mh_kernel.cuh — K2: fused multi-step Metropolis–Hastings kernel (sm_100a).
//
// Replaces, as ONE kernel per run, the per-chain host loop of
//   MHMarkovChain::step                      /root/reference/src/metropolis_hastings.rs:306-318
//   IsotropicGaussian::{sample, logp}        /root/reference/src/distributions.rs:368-390
//   Gaussian2D / IsotropicGaussian targets   /root/reference/src/distributions.rs:195-207, 398-406
//   run_chain + ChainRunner::run stack       /root/reference/src/core.rs:95-115, 219-229
//
// Mapping: one thread per chain; state, proposal and log densities live in registers for every
// step of the launch.  The f64 [chains, samples, dim] write-out (Trace -> f64, core.rs:34-51) is
// staged in shared memory for kStageDoubles/d steps and flushed row-by-row, so each chain writes
// 256 contiguous, 256-byte aligned bytes per flush instead of d*8 bytes at a 8*n*d stride.
//
// Compiled twice like K1: GM_EXACT=0 (FMA contraction, log-ratio without the cancelling proposal
// terms) and GM_EXACT=1 with --fmad=false (reference operation order: bit-for-bit the CPU oracle
// given the same normals and ln u).  In both modes the current log density is carried over from
// the previous step instead of being recomputed (metropolis_hastings.rs:308 recomputes it; the
// value is the same function of the same state, hence bit-identical).
#pragma once
#include "kernels.h"
#include "philox.cuh"

#include <cmath>
#include <type_traits>

#ifndef GM_EXACT
#define GM_EXACT 0
#endif
#if GM_EXACT
#define GM_NS exact
#else
#define GM_NS fast
#endif

namespace gm {
namespace GM_NS {

constexpr bool kMhExact = (GM_EXACT != 0);
constexpr int kMhBlock = 128;
#ifndef GM_MH_STAGE
#define GM_MH_STAGE 32
#endif
constexpr int kStageDoubles = GM_MH_STAGE;   // doubles staged per chain between flushes (256 B)
constexpr int kStageStride = kStageDoubles + 1;  // +1: conflict-free column writes

template <class T>
struct MhArgs {
  int kind;
  int d;
  T sp[kMaxScalarParams];
  T inv_cov[2][2];
  T norm_const;
  T prop_std;
  T prop_logq_const;   // -d * 0.5 * ln(var * pi * std * std), host-computed (distributions.rs:388)
  size_t n_chains;
  unsigned long long chain_offset;
  PhiloxKey key;
  PhiloxRoundKeys rk;   // the ten round keys of `key`, computed on the host: read as constant-bank operands
  uint32_t step_base, n_steps, n_skip;
  const T* dparams;     // device parameter block of a plugin target (gmcmc_custom_target.cuh), else null
  T* state;
  double* out;
  size_t out_n;
  uint32_t out_t0;
  unsigned long long* accept_total;
  const T* inj_normals;
  const T* inj_lnu;
  T* diag_logratio;
  uint8_t* diag_acc;
  float* diag_draws;    // [n, C, 3] (z0, z1, accept uniform) actually used by mh_run2_kernel<.., REC = true>, else null
};

// Target::unnorm_logp for the MH path (distributions.rs:107-110): log density only.
// Loop invariants of a target, evaluated once per thread with the reference's own expressions (so the
// values are the ones distributions.rs:195-207 recomputes on every call).
template <class T>
struct MhInv { T c0, c1, c2, c3; };

template <class T, int KIND>
__device__ __forceinline__ MhInv<T> mh_prepare(const MhArgs<T>& a) {
  MhInv<T> v{T(0), T(0), T(0), T(0)};
  if constexpr (KIND == 0) {
    v.c0 = a.sp[0] * a.sp[0];
  } else if constexpr (KIND == 1) {
    const T ca = a.sp[2], cb = a.sp[3], cc = a.sp[4], cd = a.sp[5];
    const T det = ca * cd - cb * cc;
    v.c0 = cd / det; v.c1 = (-cb) / det; v.c2 = (-cc) / det; v.c3 = ca / det;
  }
  return v;
}

template <class T, int MAXD, int KIND, bool FULL>
__device__ __forceinline__ T mh_logp(const MhArgs<T>& a, const MhInv<T>& inv, const T (&x)[MAXD]) {
  const int d = FULL ? MAXD : a.d;
  switch (KIND) {
    case 0: {  // IsotropicGaussian, distributions.rs:398-406
      T sum = T(0);
#pragma unroll
      for (int i = 0; i < MAXD; ++i)
        if (i < d) sum = sum + x[i] * x[i];
      return -T(0.5) * sum / inv.c0;
    }
    case 1: {  // Gaussian2D, distributions.rs:195-207
      const T d0 = x[0] - a.sp[0], d1 = x[MAXD > 1 ? 1 : 0] - a.sp[1];
      const T i00 = inv.c0, i01 = inv.c1, i10 = inv.c2, i11 = inv.c3;
      const T v0 = d0 * i00 + d1 * i10;
      const T v1 = d0 * i01 + d1 * i11;
      return -T(0.5) * (v0 * d0 + v1 * d1);
    }
    case 2: {  // DiffableGaussian2D, distributions.rs:265-291
      const T d0 = x[0] - a.sp[0], d1 = x[MAXD > 1 ? 1 : 0] - a.sp[1];
      const T z0 = d0 * a.inv_cov[0][0] + d1 * a.inv_cov[1][0];
      const T z1 = d0 * a.inv_cov[0][1] + d1 * a.inv_cov[1][1];
      const T quad = z0 * d0 + z1 * d1;
      return a.norm_const - quad * T(0.5);
    }
    case 4: {  // Rosenbrock2D, distributions.rs:502-515
      const T u = (-x[0]) + a.sp[0];
      const T t = x[MAXD > 1 ? 1 : 0] - x[0] * x[0];
      return -(u * u + (t * t) * a.sp[1]);
    }
    case 5: {  // RosenbrockND, distributions.rs:544-554
      T s = T(0);
#pragma unroll
      for (int i = 0; i < MAXD - 1; ++i) {
        if (i < d - 1) {
          const T t = x[i + 1] - x[i] * x[i];
          const T u = (-x[i]) + T(1);
          s = s + ((t * t) * T(100) + u * u);
        }
      }
      return -s;
    }
  }
  return T(0);
}

// IsotropicGaussian::logp(from, to), distributions.rs:378-390
template <class T, int MAXD, bool FULL>
__device__ __forceinline__ T mh_logq(const MhArgs<T>& a, const T (&from)[MAXD], const T (&to)[MAXD]) {
  T lp = T(0);
  const T var = a.prop_std * a.prop_std;
#pragma unroll
  for (int i = 0; i < MAXD; ++i) {
    if (FULL || i < a.d) {
      const T diff = to[i] - from[i];
      const T exponent = -(diff * diff) / (T(2) * var);
      lp = lp + exponent;
    }
  }
  return lp + a.prop_logq_const;
}

#ifndef GM_MH_MINB
#define GM_MH_MINB 6
#endif

// One launch = n_skip discarded transitions, then (n_steps - n_skip) recorded ones.  INSTR = the
// instrumented variant used by the parity tests (injected normals / ln u, per-step diagnostics).
// LP = void: the built-in target KIND (mh_logp above).  Otherwise LP::logp<T>(x, params) is a plugin target's log density
// (gmcmc_custom_target.cuh instantiates this kernel for it).
template <class T, int MAXD, int KIND, bool FULL, bool INSTR, class LP = void>
__global__ void __launch_bounds__(kMhBlock, GM_MH_MINB) mh_run_kernel(const MhArgs<T> a) {
  __shared__ double stage[kMhBlock * kStageStride];

  const size_t chain = (size_t)blockIdx.x * kMhBlock + threadIdx.x;
  const bool active = chain < a.n_chains;
  const unsigned long long gchain = a.chain_offset + chain;
  const int d = FULL ? MAXD : a.d;
  const int lane = threadIdx.x & 31;
  const int warp_row0 = threadIdx.x & ~31;
  const size_t warp_first_chain = (size_t)blockIdx.x * kMhBlock + warp_row0;
  const int steps_per_flush = kStageDoubles / d;  // >= 1 (d <= kStageDoubles)
  // rows of this warp that exist (the last warp of the grid may be ragged)
  const int warp_rows = a.n_chains > warp_first_chain
                            ? (int)(a.n_chains - warp_first_chain < 32 ? a.n_chains - warp_first_chain : 32) : 0;
  const size_t row_pitch = a.out_n * (size_t)d;                       // doubles between consecutive chains
  double* out_lane = a.out ? a.out + (warp_first_chain * a.out_n + a.out_t0) * (size_t)d + lane : nullptr;
  const double* stage_lane = stage + warp_row0 * kStageStride + lane;
  double* const stage_row = stage + threadIdx.x * kStageStride;
  const PhiloxRoundKeys& rk = a.rk;   // host-computed round keys: constant-bank operands

  T x[MAXD];
#pragma unroll
  for (int i = 0; i < MAXD; ++i) x[i] = (active && (FULL || i < d)) ? a.state[chain * d + i] : T(0);
  const MhInv<T> inv = mh_prepare<T, KIND>(a);
  auto logp_of = [&](const T (&v)[MAXD]) -> T {
    if constexpr (std::is_void<LP>::value) return mh_logp<T, MAXD, KIND, FULL>(a, inv, v);
    else return LP::template logp<T>(v, a.dparams);
  };
  T lp_cur = logp_of(x);

  unsigned int n_accept = 0;
  int staged = 0;                 // doubles currently in this chain's stage row

  // one transition; returns nothing, updates x / lp_cur / n_accept
  auto transition = [&](const uint32_t s) {
    const uint32_t step = a.step_base + s;
    // ---- proposal noise (distributions.rs:368-376): stream 0 of the RNG contract
    T z[MAXD];
    uint4 blk0 = make_uint4(0u, 0u, 0u, 0u);
    if (INSTR && a.inj_normals) {
#pragma unroll
      for (int i = 0; i < MAXD; ++i)
        z[i] = (active && (FULL || i < d)) ? a.inj_normals[((size_t)s * a.n_chains + chain) * d + i] : T(0);
    } else if constexpr (!kMhExact) {
      // fast mode (both dtypes): 24-bit uniforms + MUFU Box-Muller, 4 normals per block.  For d <= 2 the
      // second half of block 0 is left for the accept uniform (one Philox block per transition).
#pragma unroll
      for (int b = 0; b < (MAXD + 3) / 4; ++b) {
        if (FULL || b * 4 < d) {
          const uint4 r = philox4x32_10(philox_ctr(gchain, step, 0u, (uint32_t)b), rk);
          if (b == 0) blk0 = r;
          float zb[4];
          if constexpr (MAXD <= 2) normals_pair_fast(r.x, r.y, zb[0], zb[1]);
          else normals_from_block<false>(r, zb);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (b * 4 + k < MAXD) z[b * 4 + k] = (T)zb[k];
        }
      }
    } else {
      constexpr int NPB = NormalsPerBlock<T>::value;
#pragma unroll
      for (int b = 0; b < (MAXD + NPB - 1) / NPB; ++b) {
        if (FULL || b * NPB < d) {
          T zb[NPB];
          normals_from_block<true>(philox4x32_10(philox_ctr(gchain, step, 0u, (uint32_t)b), rk), zb);
#pragma unroll
          for (int k = 0; k < NPB; ++k)
            if (b * NPB + k < MAXD) z[b * NPB + k] = zb[k];
        }
      }
    }
    T xp[MAXD];
#pragma unroll
    for (int i = 0; i < MAXD; ++i) xp[i] = (FULL || i < d) ? (x[i] + z[i] * a.prop_std) : T(0);

    // ---- log acceptance ratio (metropolis_hastings.rs:308-312)
    const T lp_prop = logp_of(xp);
    T log_ratio;
    if constexpr (kMhExact) {
      const T q_fwd = mh_logq<T, MAXD, FULL>(a, x, xp);
      const T q_bwd = mh_logq<T, MAXD, FULL>(a, xp, x);
      log_ratio = (lp_prop + q_bwd) - (lp_cur + q_fwd);
    } else {
      log_ratio = lp_prop - lp_cur;  // symmetric proposal: q_fwd == q_bwd bit-for-bit
    }
    // ---- accept iff log_ratio > ln u (strict; metropolis_hastings.rs:313-316)
    bool accept;
    if (INSTR && a.inj_lnu) {
      const T ln_u = active ? a.inj_lnu[(size_t)s * a.n_chains + chain] : T(0);
      accept = log_ratio > ln_u;
    } else if constexpr (kMhExact) {
      const T ln_u = log(accept_uniform<T>(philox4x32_10(philox_ctr(gchain, step, 1u, 0u), rk)));
      accept = log_ratio > ln_u;
    } else {
      // fast mode: the uniform comes from words 2-3 of block 0 when d <= 2, else from stream 1.  The f32
      // MUFU logarithm of its leading 24 bits decides unless the margin is inside the error bound; only
      // then is the full-width uniform built and its logarithm evaluated in T, so the decision is always
      // the one the full-precision comparison gives.
      uint4 r = blk0;
      if (MAXD > 2) { r = philox4x32_10(philox_ctr(gchain, step, 1u, 0u), rk); r.z = r.x; r.w = r.y; }
      float l2;
      asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u01(r.z)));
      const float ln_u32 = l2 * 0.6931471805599453f;
      const float lr32 = (float)log_ratio;
      const float diff = lr32 - ln_u32;
      // error bound of diff: 2^-24 truncation of u + ~1e-6 (1 + |ln u|) MUFU + 1e-7 |log_ratio|
      if (fabsf(diff) > 2e-5f * (1.0f + fabsf(ln_u32) + fabsf(lr32))) {
        accept = diff > 0.0f;
      } else {
        T u;
        if constexpr (sizeof(T) == 8) u = u01d(r.z, r.w); else u = u01(r.z);
        accept = log_ratio > log(u);
      }
    }
#pragma unroll
    for (int i = 0; i < MAXD; ++i) x[i] = accept ? xp[i] : x[i];
    lp_cur = accept ? lp_prop : lp_cur;
    n_accept += (accept && active) ? 1u : 0u;
    if (INSTR && a.diag_logratio && active) {
      a.diag_logratio[(size_t)s * a.n_chains + chain] = log_ratio;
      a.diag_acc[(size_t)s * a.n_chains + chain] = accept ? 1 : 0;
    }
  };

  // stage -> [chain, slot .. slot + staged/d, :] for the warp's rows (256 contiguous bytes per chain)
  auto flush = [&]() {
    __syncwarp();
    if (lane < staged) {
      double* dst = out_lane;
      const double* src = stage_lane;
      if (warp_rows == 32) {
#pragma unroll 8
        for (int r = 0; r < 32; ++r) { __stcs(dst, *src); dst += row_pitch; src += kStageStride; }
      } else {
        for (int r = 0; r < warp_rows; ++r) { __stcs(dst, *src); dst += row_pitch; src += kStageStride; }
      }
    }
    __syncwarp();
    out_lane += staged;
    staged = 0;
  };

  uint32_t s = 0;
  const uint32_t n_skip = a.n_skip < a.n_steps ? a.n_skip : a.n_steps;
  for (; s < n_skip; ++s) transition(s);                 // burn-in: nothing recorded
  if (out_lane) {
    const int flush_at = steps_per_flush * d;
    for (; s < a.n_steps; ++s) {
      transition(s);
#pragma unroll
      for (int i = 0; i < MAXD; ++i)
        if (FULL || i < d) stage_row[staged + i] = (double)x[i];
      staged += d;
      if (staged == flush_at) flush();
    }
    if (staged) flush();
  } else {
    for (; s < a.n_steps; ++s) transition(s);
  }

  if (active) {
#pragma unroll
    for (int i = 0; i < MAXD; ++i)
      if (FULL || i < d) a.state[chain * d + i] = x[i];
  }
  for (int o = 16; o > 0; o >>= 1) n_accept += __shfl_xor_sync(0xffffffffu, n_accept, o);
  if (lane == 0 && n_accept) atomicAdd(a.accept_total, (unsigned long long)n_accept);
}

// ------------------------------------------------------------------------------------------------
// K2 fast path: d == 2 (every 2-D target; BASELINE config 2), fast math mode, production (no injected
// draws).  Same transition as mh_run_kernel; what differs is how the per-step cost is kept near the 16
// bytes the step writes:
//   * one Philox block per TWO transitions (20-bit radius / angle uniforms, a 23-bit accept uniform: 64 bits per
//     transition), the inputs of transition s + 1 produced while transition s runs its f64 dependency chain;
//   * uniforms are built without int->float conversions: bits become the mantissa of a float in [1,2);
//   * the accept test runs in the linear domain: e = ex2.approx(log_ratio * log2 e) against the uniform, and only
//     when |e - u| is inside the error bound of e (MUFU + argument rounding) is `log_ratio > ln u` evaluated in T
//     with the same uniform — the decision is always the full-precision one (metropolis_hastings.rs:313-316);
//   * the f64 sample of a step is ONE 16-byte unit: it is staged with st.shared.v2 under an XOR swizzle
//     (conflict-free for the per-chain writes and for the per-row reads) and flushed every GM_MH2_STEPS
//     steps as LDS.128 + STG.128 pairs, 256 (16 steps) or 512 (32 steps) contiguous bytes per chain.
// Measured on B200 (tools/mh_bench.cu, tools/microbench_write.cu): a write-only stream of 256-byte pieces
// reaches 5.3 TB/s, of 512-byte pieces 6.6 TB/s, cudaMemset 7.4 TB/s; the kernel writes 4.6 TB/s (3.65 ms per
// 1000 steps of 1,048,576 chains, 70 % of the measured copy bandwidth) with either staging depth and any occupancy
// from 12 to 32 warps per SM: it is bound by instruction dispatch (89 warp-instructions per step, most of them on
// half-rate pipes: Philox's LOP3 / IMAD.WIDE, 13 FP64, selects), not by HBM — profiles/r1_mh_run2_kernel_full.txt.
// ------------------------------------------------------------------------------------------------
#ifndef GM_MH2_STEPS
#define GM_MH2_STEPS 16
#endif
#ifndef GM_MH2_MINB
#define GM_MH2_MINB 6
#endif
#ifndef GM_MH2_UNROLL
#define GM_MH2_UNROLL 8
#endif
constexpr int kMh2Steps = GM_MH2_STEPS;   // steps staged per chain between flushes: 16 (256 B per chain) or 32 (512 B)
constexpr int kMh2Unroll = GM_MH2_UNROLL;
constexpr int kMh2Row = kMh2Steps * 16;   // bytes of one chain's stage row
constexpr size_t kMh2Smem = (size_t)kMhBlock * kMh2Row;
static_assert(kMh2Steps == 16 || kMh2Steps == 32, "the flush maps 32 lanes onto 16-byte units of one or two chains");


// REC = true: the same kernel with per-step stores of what it drew and decided (proposal noise, accept uniform, log
// ratio, decision) for the per-step parity test against the oracle (gmcmc_mh_record); no arithmetic differs.
template <class T, int KIND, bool REC = false>
__global__ void __launch_bounds__(kMhBlock, GM_MH2_MINB) mh_run2_kernel(const MhArgs<T> a) {
  extern __shared__ __align__(1024) unsigned char stage[];   // kMhBlock rows of kMh2Row bytes (row-aligned: the flush XORs address bits)

  const size_t chain = (size_t)blockIdx.x * kMhBlock + threadIdx.x;
  const bool active = chain < a.n_chains;
  const unsigned long long gchain = a.chain_offset + chain;
  const int lane = threadIdx.x & 31;
  const int warp_row0 = threadIdx.x & ~31;
  const size_t warp_first_chain = (size_t)blockIdx.x * kMhBlock + warp_row0;
  const int warp_rows = a.n_chains > warp_first_chain
                            ? (int)(a.n_chains - warp_first_chain < 32 ? a.n_chains - warp_first_chain : 32) : 0;
  const PhiloxRoundKeys& rk = a.rk;

  // shared-memory addresses (32-bit).  Unit (chain c, step t) of a warp lives at row c, 16-byte column t ^ (c & (S-1)):
  // the 8 lanes of a quarter-warp then hit 8 different bank groups both when every chain writes its step-t unit and
  // when a row is read back along t.
  constexpr int kColMask = kMh2Steps - 1;
  const uint32_t stage_base = (uint32_t)__cvta_generic_to_shared(stage) + (uint32_t)warp_row0 * kMh2Row;
  uint32_t st_addr = stage_base + (uint32_t)lane * kMh2Row + ((uint32_t)(lane & kColMask) << 4);
  asm volatile("mov.b32 %0, %0;" : "+r"(st_addr));   // opaque: keep it in a register instead of rematerialising it every step
  // flush: one store instruction moves 512 contiguous bytes per chain (32 steps) or 2 x 256 (16 steps)
  constexpr int kCpi = 32 / kMh2Steps;               // chains per flush instruction
  const int f_sub = lane / kMh2Steps, f_t = lane & kColMask;

  T x[2];
  x[0] = active ? a.state[chain * 2 + 0] : T(0);
  x[1] = active ? a.state[chain * 2 + 1] : T(0);
  const MhInv<T> inv = mh_prepare<T, KIND>(a);
  T lp_cur = mh_logp<T, 2, KIND, true>(a, inv, x);
  unsigned int n_accept = 0;

  // Random inputs.  ONE Philox block feeds TWO transitions (Philox4x32-10 is 40 of the kernel's ~100 instructions per
  // step): transition t (absolute index) takes words (0, 1) of block t >> 1 when t is even, words (2, 3) when odd.  Of a
  // word pair (w0, w1): bits 31..12 of w0 -> radius uniform (k + 1/2) 2^-20, bits 31..12 of w1 -> angle
  // 2 pi (k + 1/2) 2^-20 - pi (a grid that maps onto itself under z -> -z, so the proposal stays exactly symmetric),
  // the low 12 + 11 bits -> the accept uniform (k + 1/2) 2^-23, which IS the uniform of this mode (no hidden bits).
  // The inputs of transition s + 1 are produced while transition s runs its state-dependent f64 chain.
  float z0, z1, ua;
  uint32_t sv0 = 0u, sv1 = 0u;      // words (2, 3) of the current block: the odd transition's pair
  auto derive = [&](const uint32_t w0, const uint32_t w1) {
    // proposal noise (distributions.rs:368-376): Box-Muller
    const float u_rad = __uint_as_float(((w0 >> 9) & 0x007ffff8u) | 0x3f800000u) - 0.999999523162841796875f;   // f - (1 - 2^-21)
    const float ang = fmaf(__uint_as_float(((w1 >> 9) & 0x007ffff8u) | 0x3f800000u), 6.283185307179586f, -9.424774964f);  // -3 pi + pi 2^-20
    float rad, sn, cs;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(u_rad));
    rad *= -1.3862943611198906f;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(rad));
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(ang));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(ang));
    z0 = rad * cs; z1 = rad * sn;
    ua = __uint_as_float(((w0 & 0xfffu) << 11) | (w1 & 0x7ffu) | 0x3f800000u) - 0.99999994039535522461f;     // (k + 1/2) 2^-23
  };
  // MODE 1: t is even (new block); 0: t is odd (saved words); 2: decided at run time (burn-in and tail loops)
  auto draw = [&](auto mode, const uint32_t t) {
    constexpr int MODE = decltype(mode)::value;
    uint32_t w0 = sv0, w1 = sv1;
    if (MODE == 1 || (MODE == 2 && (t & 1u) == 0u)) {
      const uint4 r = philox4x32_10(philox_ctr(gchain, t >> 1, 0u, 0u), rk);
      w0 = r.x; w1 = r.y; sv0 = r.z; sv1 = r.w;
    }
    derive(w0, w1);
  };
  using Saved = std::integral_constant<int, 0>;
  using Fresh = std::integral_constant<int, 1>;
  using Either = std::integral_constant<int, 2>;
  {
    const uint4 r = philox4x32_10(philox_ctr(gchain, a.step_base >> 1, 0u, 0u), rk);
    sv0 = r.z; sv1 = r.w;
    if (a.step_base & 1u) derive(r.z, r.w); else derive(r.x, r.y);
  }

  auto transition = [&](const uint32_t s, auto next_mode) {     // next_mode: how transition s + 1 gets its word pair
    const float c0 = z0, c1 = z1, cu = ua;
    draw(next_mode, a.step_base + s + 1u);
    T xp[2];
    xp[0] = x[0] + (T)c0 * a.prop_std;
    xp[1] = x[1] + (T)c1 * a.prop_std;
    // ---- log acceptance ratio (metropolis_hastings.rs:308-312); symmetric proposal: q_fwd == q_bwd bit-for-bit
    const T lp_prop = mh_logp<T, 2, KIND, true>(a, inv, xp);
    const T log_ratio = lp_prop - lp_cur;
    // ---- accept iff log_ratio > ln u (strict; metropolis_hastings.rs:313-316).  Linear-domain test first; inside the
    // error bound of e (MUFU + rounding of its argument) the comparison is redone in T with the very same uniform.
    const float lr32 = (float)log_ratio;
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(lr32 * 1.4426950408889634f));
    const float diff = e - cu;
    const float tol = e * (1.0f + fabsf(lr32)) * 1e-6f;
    bool accept;
    if (fabsf(diff) > tol) {
      accept = diff > 0.0f;
    } else {
      accept = log_ratio > log((T)cu);
    }
    x[0] = accept ? xp[0] : x[0];
    x[1] = accept ? xp[1] : x[1];
    lp_cur = accept ? lp_prop : lp_cur;
    n_accept += accept ? 1u : 0u;
    if constexpr (REC) {
      if (active) {
        const size_t idx = (size_t)s * a.n_chains + chain;
        a.diag_draws[idx * 3 + 0] = c0; a.diag_draws[idx * 3 + 1] = c1; a.diag_draws[idx * 3 + 2] = cu;
        a.diag_logratio[idx] = log_ratio;
        a.diag_acc[idx] = accept ? 1 : 0;
      }
    }
  };

  uint32_t s = 0;
  const uint32_t n_skip = a.n_skip < a.n_steps ? a.n_skip : a.n_steps;
  for (; s < n_skip; ++s) transition(s, Either{});       // burn-in: nothing recorded
  if (a.out) {
    // this lane's flush destination: [chain = warp_first + f_sub (+kCpi per iteration), slot = out_t0 + f_t (+S per flush)]
    double* out_lane = a.out + ((warp_first_chain + f_sub) * a.out_n + a.out_t0 + f_t) * 2;
    const size_t pitch = a.out_n * 2 * kCpi;             // doubles between the chains of consecutive iterations
    auto flush = [&](const int cnt) {
      __syncwarp();
      if (f_t < cnt) {
        double* dst = out_lane;
        if (warp_rows == 32) {
          // chain c = kCpi i + f_sub: c & mask = ((kCpi i) & mask) ^ f_sub (f_sub only occupies the bit kCpi i leaves free), so
          // the swizzled address is one XOR of a per-lane base with a compile-time constant, plus a compile-time row offset
          const uint32_t lane_base = stage_base + (uint32_t)f_sub * kMh2Row + ((uint32_t)(f_t ^ f_sub) << 4);
#pragma unroll
          for (int i = 0; i < 32 / kCpi; ++i) {
            const uint32_t addr = (lane_base ^ ((uint32_t)((kCpi * i) & kColMask) << 4)) + (uint32_t)(kCpi * i) * kMh2Row;
            double v0, v1;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(addr));
            __stcs(reinterpret_cast<double2*>(dst), make_double2(v0, v1));
            dst += pitch;
          }
        } else {
          for (int i = 0; i < 32 / kCpi; ++i) {
            const int c = kCpi * i + f_sub;
            if (c < warp_rows) {
              const uint32_t addr = stage_base + (uint32_t)c * kMh2Row + ((uint32_t)(f_t ^ (c & kColMask)) << 4);
              double v0, v1;
              asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(addr));
              __stcs(reinterpret_cast<double2*>(dst), make_double2(v0, v1));
            }
            dst += pitch;
          }
        }
      }
      __syncwarp();
      out_lane += kMh2Steps * 2;
    };
    auto record = [&](const int t) {
      asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(st_addr ^ ((uint32_t)t << 4)), "d"((double)x[0]), "d"((double)x[1]) : "memory");
    };
    // groups of kMh2Steps recorded transitions; the parity of the absolute index is the same at every group start, so
    // inside a group it is a compile-time fact which transitions open a new Philox block
    const bool start_even = ((a.step_base + s) & 1u) == 0u;
    while (a.n_steps - s >= (uint32_t)kMh2Steps) {
      if (start_even) {
#pragma unroll kMh2Unroll / 2
        for (int t = 0; t < kMh2Steps; t += 2) { transition(s + t, Saved{}); record(t); transition(s + t + 1, Fresh{}); record(t + 1); }
      } else {
#pragma unroll kMh2Unroll / 2
        for (int t = 0; t < kMh2Steps; t += 2) { transition(s + t, Fresh{}); record(t); transition(s + t + 1, Saved{}); record(t + 1); }
      }
      s += kMh2Steps;
      flush(kMh2Steps);
    }
    if (s < a.n_steps) {
      int t = 0;
      for (; s < a.n_steps; ++s, ++t) { transition(s, Either{}); record(t); }
      flush(t);
    }
  } else {
    for (; s < a.n_steps; ++s) transition(s, Either{});
  }

  if (active) { a.state[chain * 2 + 0] = x[0]; a.state[chain * 2 + 1] = x[1]; }
  if (!active) n_accept = 0;
  for (int o = 16; o > 0; o >>= 1) n_accept += __shfl_xor_sync(0xffffffffu, n_accept, o);
  if (lane == 0 && n_accept) atomicAdd(a.accept_total, (unsigned long long)n_accept);
}

template <class T, int KIND, bool REC = false>
inline cudaError_t mh_launch_run2(const MhArgs<T>& a, unsigned blocks, cudaStream_t st) {
  auto kern = mh_run2_kernel<T, KIND, REC>;
  if (kMh2Smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMh2Smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<blocks, kMhBlock, kMh2Smem, st>>>(a);
  return cudaGetLastError();
}

template <class T>
inline MhArgs<T> make_mh_args(const MhLaunch& L) {
  MhArgs<T> a;
  a.kind = L.tgt.kind;
  a.d = L.tgt.dim;
  for (int i = 0; i < kMaxScalarParams; ++i) a.sp[i] = (T)L.tgt.sp[i];
  a.inv_cov[0][0] = a.inv_cov[0][1] = a.inv_cov[1][0] = a.inv_cov[1][1] = T(0);
  a.norm_const = T(0);
  if (L.tgt.kind == 2) {
    // DiffableGaussian2D::new (distributions.rs:229-253), evaluated in T on the host
    volatile T c00 = a.sp[2], c01 = a.sp[3], c10 = a.sp[4], c11 = a.sp[5];
    volatile T m1 = c00 * c11;
    volatile T m2 = c01 * c10;
    volatile T det_cov = m1 - m2;
    volatile T inv_det = T(1) / det_cov;
    a.inv_cov[0][0] = c11 * inv_det;
    a.inv_cov[0][1] = -c01 * inv_det;
    a.inv_cov[1][0] = -c10 * inv_det;
    a.inv_cov[1][1] = c00 * inv_det;
    T logdet = std::log((T)det_cov);
    T two = T(1) + T(1);
    const T pi = (T)3.14159265358979323846264338327950288;
    volatile T l2pi = std::log(two * pi);
    volatile T tl = two * l2pi;
    volatile T sum = tl + logdet;
    a.norm_const = -sum / two;
  }
  a.prop_std = (T)L.prop_std;
  {
    // distributions.rs:388, in T with glibc log (same as the oracle)
    const T pi = (T)3.14159265358979323846264338327950288;
    volatile T var = a.prop_std * a.prop_std;
    volatile T t1 = var * pi;
    volatile T t2 = t1 * a.prop_std;
    volatile T t3 = t2 * a.prop_std;
    volatile T lg = std::log((T)t3);
    volatile T nd = -(T)L.tgt.dim;
    volatile T h = nd * T(0.5);
    a.prop_logq_const = h * lg;
  }
  a.dparams = (const T*)L.tgt.dparams;
  a.n_chains = L.n_chains;
  a.chain_offset = L.chain_offset;
  a.key = PhiloxKey{(uint32_t)L.seed, (uint32_t)(L.seed >> 32)};
  a.rk = philox_round_keys(a.key);
  a.step_base = L.step_base; a.n_steps = L.n_steps; a.n_skip = L.n_skip;
  a.state = (T*)L.state;
  a.out = L.out; a.out_n = L.out_n; a.out_t0 = L.out_t0;
  a.accept_total = L.accept_total;
  a.inj_normals = (const T*)L.inj_normals; a.inj_lnu = (const T*)L.inj_lnu;
  a.diag_logratio = (T*)L.diag_logratio; a.diag_acc = L.diag_acc; a.diag_draws = L.diag_draws;
  return a;
}

template <class T, int MAXD, int KIND, bool FULL>
inline cudaError_t mh_launch_one(const MhLaunch& L, cudaStream_t st) {
  MhArgs<T> a = make_mh_args<T>(L);
  const unsigned blocks = (unsigned)((L.n_chains + kMhBlock - 1) / kMhBlock);
  if (a.diag_draws) {   // gmcmc_mh_record: the production 2-D fast kernel, instrumented
    if constexpr (!kMhExact && MAXD == 2 && FULL) return mh_launch_run2<T, KIND, true>(a, blocks, st);
    else return cudaErrorInvalidValue;
  } else if (a.inj_normals || a.inj_lnu || a.diag_logratio) {
    mh_run_kernel<T, MAXD, KIND, FULL, true><<<blocks, kMhBlock, 0, st>>>(a);
  } else {
    if constexpr (!kMhExact && MAXD == 2 && FULL) return mh_launch_run2<T, KIND>(a, blocks, st);
    else mh_run_kernel<T, MAXD, KIND, FULL, false><<<blocks, kMhBlock, 0, st>>>(a);
  }
  return cudaGetLastError();
}

template <class T, int MAXD, int KIND>
inline cudaError_t mh_launch_full(const MhLaunch& L, cudaStream_t st) {
  if (L.tgt.dim == MAXD) return mh_launch_one<T, MAXD, KIND, true>(L, st);
  return mh_launch_one<T, MAXD, KIND, false>(L, st);
}

template <class T, int KIND>
inline cudaError_t mh_dispatch_dim(const MhLaunch& L, cudaStream_t st) {
  const int d = L.tgt.dim;
  if (d < 1 || d > kStageDoubles) return cudaErrorInvalidValue;
  if (d <= 2) return mh_launch_full<T, 2, KIND>(L, st);
  if (d <= 4) return mh_launch_full<T, 4, KIND>(L, st);
  if (d <= 8) return mh_launch_full<T, 8, KIND>(L, st);
  if (d <= 16) return mh_launch_full<T, 16, KIND>(L, st);
  return mh_launch_full<T, 32, KIND>(L, st);
}

template <class T>
inline cudaError_t mh_dispatch(const MhLaunch& L, cudaStream_t st) {
  switch (L.tgt.kind) {
    case 0: return mh_dispatch_dim<T, 0>(L, st);
    case 5: return mh_dispatch_dim<T, 5>(L, st);
    case 1: return L.tgt.dim == 2 ? mh_launch_one<T, 2, 1, true>(L, st) : cudaErrorInvalidValue;
    case 2: return L.tgt.dim == 2 ? mh_launch_one<T, 2, 2, true>(L, st) : cudaErrorInvalidValue;
    case 4: return L.tgt.dim == 2 ? mh_launch_one<T, 2, 4, true>(L, st) : cudaErrorInvalidValue;
  }
  return cudaErrorInvalidValue;
}

}  // namespace GM_NS

#ifndef GM_MH_NO_BUILTIN_LAUNCHERS   // plugins include this header for the kernel template only
#if GM_EXACT
cudaError_t launch_mh_exact(const MhLaunch& L, cudaStream_t st) {
  return L.tgt.dtype == 0 ? exact::mh_dispatch<float>(L, st) : exact::mh_dispatch<double>(L, st);
}
#else
cudaError_t launch_mh_fast(const MhLaunch& L, cudaStream_t st) {
  return L.tgt.dtype == 0 ? fast::mh_dispatch<float>(L, st) : fast::mh_dispatch<double>(L, st);
}
#endif
#endif  // GM_MH_NO_BUILTIN_LAUNCHERS

}  // namespace gm
