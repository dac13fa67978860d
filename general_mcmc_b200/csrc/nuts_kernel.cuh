// nuts_kernel.cuh — K5: No-U-Turn sampler, one launch per run segment (sm_100a).
//
// Replaces the host-driven recursion of
//   GenericNUTSChain::step                      /root/reference/src/generic_nuts.rs:755-925
//   build_tree_with_mass (13-tuple recursion)   /root/reference/src/generic_nuts.rs:1153-1341
//   leapfrog_with_mass / stop_criterion         /root/reference/src/generic_nuts.rs:1396-1418, 1357-1378
//   find_reasonable_epsilon_with_mass           /root/reference/src/generic_nuts.rs:1025-1102
//   init_chain_state / run                      /root/reference/src/generic_nuts.rs:667-753
//
// The recursion becomes an iterative binary-counter tree.  Leaf c of a depth-j subtree build merges
// with the pending left subtree of every level k whose bit is set in c, in increasing k — exactly the
// order in which the reference's recursion returns, so the per-node uniforms are consumed in the same
// (post-order) sequence.  A failed leaf / merge (s' = false) keeps merging through ALL set bits (the
// reference's ancestors whose left child was valid still merge; ancestors reached through a left child
// pass the result up untouched).  Only two kinds of vectors are checkpointed:
//   first[slot]  (q, p) of every even leaf, slot = popcount(c >> 1): the first leaf of every pending
//                subtree (U-turn test of a merge = first leaf of the left subtree vs. the current leaf)
//   prime[k]     proposal of the pending subtree at level k >= 1 (level 0: its proposal is its own leaf)
// alpha' and n_alpha' are plain sums over the leaves of a build, n' is carried per level.
//
// Mapping: as K1, a chain is `lpc` adjacent lanes with EPL coordinates each; q, p, grad and the
// current proposal live in registers.  Chains of one warp progress independently (different tree
// depths, different transition counts); every loop iteration evaluates ONE gradient per chain and all
// shuffles are executed by the full warp with per-chain commit masks (warp-level masking of finished
// or differently-phased chains).
#pragma once
#include "hmc_kernel.cuh"

#include <type_traits>

namespace gm {
namespace GM_NS {

constexpr int kNutsDepthCap = 20;  // safety cap when max_depth == 0 (the reference is uncapped)

template <class T>
struct NutsArgs {
  TParams<T> tp;
  size_t n_chains;
  unsigned long long chain_offset;
  PhiloxKey key;
  uint32_t step_base;     // Philox transition index of the first transition of this launch
  T* positions;           // [C, d]
  int d, d_pad, lpc;
  uint32_t n_steps;       // transitions in this launch
  uint32_t m_base;        // transitions of this run before the launch: m = m_base + s + 1 (generic_nuts.rs:756)
  uint32_t n_discard;     // dual averaging adapts while m <= n_discard (generic_nuts.rs:897)
  long long rec_off;      // transition m is recorded at slot m - rec_off when 0 <= slot < out_n
  int write_init;         // run(): sample 0 = the initial position when n_discard == 0 (nuts.rs:588-601)
  T* out;                 // [C, out_n, d] or null
  size_t out_n;
  T* eps; T* eps_bar; T* h_bar; T* mu;   // per-chain adaptation state [C]
  T target_accept;
  int max_depth;          // effective cap (1..kNutsDepthCap)
  T* ws_edges;            // [C][6][d]  q-, p-, g-, q+, p+, g+
  T* ws_first;            // [C][cap][2][d]
  T* ws_prime;            // [C][cap][d]
  int cap;
  unsigned long long* leapfrog_total;
  unsigned long long* diverge_total;
  unsigned long long* depth_total;
  unsigned long long* accept_total;   // transitions that moved the chain (a subtree proposal was accepted, :866-868)
  long long* chain_leapfrogs;   // [C] accumulated (may be null)
  // injected per-chain streams (parity tests); null -> Philox
  const double* inj_normals; size_t n_norm;
  const double* inj_exp1; size_t n_exp;
  const double* inj_unif; size_t n_unif;
  unsigned long long* inj_used;   // [C][3] consumption counters, in/out
  unsigned long long* queue;      // [1] next chain to hand out (initialised to the number of lane groups of the grid)
  // diagonal mass matrix (MassMatrix::Diagonal, generic_nuts.rs:177-304): inv = 1 / var, sqrt = sqrt(var); null = identity
  const T* mass_inv; const T* mass_sqrt;          // [C, d]
  // warm-up position statistics (RunningCov, generic_nuts.rs:81-132), updated for collect_after < m < collect_before
  T* run_mean; T* run_m2;                         // [C, d], [C, d]
  uint32_t run_n_base;                            // positions already in the statistics when this launch starts (same for every chain)
  uint32_t collect_after, collect_before;
  // dense mass matrix (MassMatrix::Dense, generic_nuts.rs:187-303): inverse mass = covariance estimate [C, d, d], lower
  // Cholesky factor of it [C, d, d], running outer-product sums of the warm-up window [C, d, d] (upper triangle)
  const T* mass_dinv; const T* mass_chol; T* run_m2d;
  int dense_active;                               // 0 until the first update: identity arithmetic (MassMatrix::Identity)
};

enum NutsPhase : int { NP_START = 0, NP_LEAF = 1, NP_END = 2, NP_DONE = 3 };

// Workspace vectors are stored lane-padded in 16-byte units (float4 / double2): lane `part` owns EPLP = roundup(EPL, 4)
// elements, unit i of lane `part` at unit index i * lpc + part — consecutive lanes are 16 bytes apart, so a quarter-warp's
// LDS.128 / STS.128 covers every bank once and the global copies are fully coalesced (3.5x fewer LSU instructions than
// element-wise accesses at EPL = 25; with the first layout, part * EPLP + i, lanes were 32 bytes apart and every shared
// access was a 2-way bank conflict: 7e8 conflict cycles per launch in profiles/r2_nuts_run_kernel_full.txt).
template <int EPL> struct Eplp { static constexpr int value = (EPL + 3) / 4 * 4; };

template <class T, int EPL>
__device__ __forceinline__ void load_slice(T (&dst)[EPL], const T* src, const Lane& ln, bool on, T fill) {
  using V = typename VecOf<T>::type;
  constexpr int VN = VecOf<T>::n;
  constexpr int EPLP = Eplp<EPL>::value;
  const V* s = reinterpret_cast<const V*>(src) + ln.part;
#pragma unroll
  for (int i = 0; i < EPLP / VN; ++i) {
    if (i * VN < EPL) {
      T e[VN];
      if (on) {
        const V v = s[i * ln.lpc];
        if constexpr (VN == 4) { e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w; }
        else { e[0] = v.x; e[1] = v.y; }
      }
#pragma unroll
      for (int k = 0; k < VN; ++k)
        if (i * VN + k < EPL) dst[i * VN + k] = (on && i * VN + k < ln.nvalid) ? e[k] : fill;
    }
  }
}
template <class T, int EPL>
__device__ __forceinline__ void store_slice(T* dst, const T (&src)[EPL], const Lane& ln, bool on) {
  using V = typename VecOf<T>::type;
  constexpr int VN = VecOf<T>::n;
  constexpr int EPLP = Eplp<EPL>::value;
  V* d = reinterpret_cast<V*>(dst) + ln.part;
  if (!on) return;
#pragma unroll
  for (int i = 0; i < EPLP / VN; ++i) {
    if (i * VN < EPL) {
      V v;
      if constexpr (VN == 4) {
        v.x = src[i * 4]; v.y = (i * 4 + 1 < EPL) ? src[i * 4 + 1 < EPL ? i * 4 + 1 : 0] : T(0);
        v.z = (i * 4 + 2 < EPL) ? src[i * 4 + 2 < EPL ? i * 4 + 2 : 0] : T(0);
        v.w = (i * 4 + 3 < EPL) ? src[i * 4 + 3 < EPL ? i * 4 + 3 : 0] : T(0);
      } else {
        v.x = src[i * 2]; v.y = (i * 2 + 1 < EPL) ? src[i * 2 + 1 < EPL ? i * 2 + 1 : 0] : T(0);
      }
      d[i * ln.lpc] = v;
    }
  }
}

// sum over the chain of f(j), j = this lane's coordinates (terms past the chain end must be exact zeros):
// the stop-criterion dot products without materialising term arrays
template <class T, int EPL, class F>
__device__ __forceinline__ T chain_sum_fn(const Lane& ln, F f) {
  if constexpr (kExact) {
    T s = T(0);
    for (int k = 0; k < ln.lpc; ++k) {
      if (ln.part == k) {
#pragma unroll
        for (int j = 0; j < EPL; ++j) s = s + f(j);
      }
      s = __shfl_sync(kFull, s, ln.gbase + k);
    }
    return s;
  } else {
    T s0 = T(0), s1 = T(0), s2 = T(0), s3 = T(0);
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      const T t = f(j);
      if ((j & 3) == 0) s0 += t; else if ((j & 3) == 1) s1 += t; else if ((j & 3) == 2) s2 += t; else s3 += t;
    }
    T s = (s0 + s1) + (s2 + s3);
    for (int o = ln.lpc >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    return s;
  }
}

// two such sums at once: in fast mode their butterfly steps interleave (half the dependent shuffle latency); exact mode
// keeps the two sequential lane-ordered sums
template <class T, int EPL, class F1, class F2>
__device__ __forceinline__ void chain_sum_fn2(const Lane& ln, F1 f1, F2 f2, T& out1, T& out2) {
  if constexpr (kExact) {
    out1 = chain_sum_fn<T, EPL>(ln, f1);
    out2 = chain_sum_fn<T, EPL>(ln, f2);
  } else {
    T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0), b0 = T(0), b1 = T(0), b2 = T(0), b3 = T(0);
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      const T t = f1(j), u = f2(j);
      if ((j & 3) == 0) { a0 += t; b0 += u; } else if ((j & 3) == 1) { a1 += t; b1 += u; }
      else if ((j & 3) == 2) { a2 += t; b2 += u; } else { a3 += t; b3 += u; }
    }
    T s1 = (a0 + a1) + (a2 + a3), s2 = (b0 + b1) + (b2 + b3);
    for (int o = ln.lpc >> 1; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(kFull, s1, o);
      s2 += __shfl_xor_sync(kFull, s2, o);
    }
    out1 = s1; out2 = s2;
  }
}

#ifndef GM_NUTS_PIN_TID
#define GM_NUTS_PIN_TID 0   // measured: 1.00e9 vs 1.02e9 leapfrogs/s with the index pinned — the rematerialisation is the cheaper choice
#endif
#ifndef GM_NUTS_MINB
// Resident CTAs per SM the kernel is register-budgeted for.  Round 1's layout gained 6 % from 4 CTAs (128 registers); with the
// hot vectors in shared memory and two leaves per pass, 3 CTAs (146 registers: the addresses and target parameters the
// 128-register build rematerialised in the loop stay in registers) are 4 % faster: 1.18e9 against 1.13e9 leapfrogs/s on
// config 5 (profiles/r2_nuts_occupancy_3_vs_4.txt).
#define GM_NUTS_MINB 3
#endif

// Targets whose padded slots (coordinates past the end of the chain in a non exact-fit decomposition) may simply hold
// zeros: a zero position / momentum there contributes exact zeros to every sum and receives a zero gradient, so the
// trajectory code needs no per-coordinate masks (which cost ~20 % of the NUTS kernel's instructions at d = 100 = 8 x 16 - 28).
template <class TAG> struct PadSafe { static constexpr bool value = false; };
template <> struct PadSafe<TagIsoGauss> { static constexpr bool value = true; };
template <> struct PadSafe<TagMixture> { static constexpr bool value = true; };

// unmasked slice moves for lane-padded vectors (padded slots hold harmless values by construction)
template <class T, int EPL>
__device__ __forceinline__ void load_slice_raw(T (&dst)[EPL], const T* src, const Lane& ln) {
  using V = typename VecOf<T>::type;
  constexpr int VN = VecOf<T>::n;
  constexpr int EPLP = Eplp<EPL>::value;
  const V* s = reinterpret_cast<const V*>(src) + ln.part;
#pragma unroll
  for (int i = 0; i < EPLP / VN; ++i) {
    if (i * VN < EPL) {
      const V v = s[i * ln.lpc];
      if constexpr (VN == 4) {
        dst[i * 4] = v.x;
        if (i * 4 + 1 < EPL) dst[i * 4 + 1 < EPL ? i * 4 + 1 : 0] = v.y;
        if (i * 4 + 2 < EPL) dst[i * 4 + 2 < EPL ? i * 4 + 2 : 0] = v.z;
        if (i * 4 + 3 < EPL) dst[i * 4 + 3 < EPL ? i * 4 + 3 : 0] = v.w;
      } else {
        dst[i * 2] = v.x;
        if (i * 2 + 1 < EPL) dst[i * 2 + 1 < EPL ? i * 2 + 1 : 0] = v.y;
      }
    }
  }
}

// Hot per-chain vectors kept in shared memory (lane-padded, wd elements each): the two trajectory edges' (q, p) — the
// other end of the whole-trajectory U-turn test and the restart point of a direction flip — and the (q, p) of the last
// even leaf, which is the first leaf of the left subtree of EVERY level-0 merge (half of all merges).  Deeper merges
// (level k >= 1: one per 2^(k+1) leaves) read their first leaf from the global workspace.
enum NutsHot : int { H_QM = 0, H_PM = 1, H_QP = 2, H_PP = 3, H_LQ = 4, H_LP = 5, H_COUNT = 6 };

// MASS: 0 = identity mass; 1 = diagonal, 2 = dense mass matrix + warm-up statistics compiled in
// (GenericNUTS::new_with_mass_matrix); the identity-mass instantiation carries none of it (measured: the run-time test
// alone cost 16 % on BASELINE config 5).
// LPC: lanes per chain as a compile-time constant (0 = run-time a.lpc): the shuffle trees unroll.
template <class T, int EPL, class TAG, bool PADDED, int MASS, int LPC>
__global__ void __launch_bounds__(kHmcBlock, GM_NUTS_MINB) nuts_run_kernel(const NutsArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  constexpr bool EP = PADDED && !PadSafe<TAG>::value;     // per-coordinate masks needed in the trajectory arithmetic
  constexpr int EPLP = Eplp<EPL>::value;
  const T pad_q = PadSafe<TAG>::value ? T(0) : T(1);      // position value of padded slots / idle lane groups

  // Lane groups are persistent workers: group `slot` starts with chain `slot` and, whenever its chain has finished
  // all its transitions, takes the next chain from a global queue — tree sizes are heavy-tailed (a few chains build
  // 20x larger trees), so a static chain -> lane-group assignment would leave most of the GPU waiting for them.
  const int lpc = LPC > 0 ? LPC : a.lpc;
  // At 125 registers the compiler rematerialises the addresses derived from the thread index inside the loop (11 % of the
  // issued instructions: profiles/r2_nuts_run_kernel_full.txt); pinning the index in a register (GM_NUTS_PIN_TID) was
  // measured 2 % slower.
  unsigned tix = threadIdx.x;
#if GM_NUTS_PIN_TID
  asm volatile("mov.b32 %0, %0;" : "+r"(tix));
#endif
  const int tid = blockIdx.x * blockDim.x + (int)tix;
  const Lane ln = make_lane<EPL>(tid, lpc, a.d);
  const size_t slot_id = (size_t)(tid / lpc);      // workspace row of this lane group
  size_t chain = slot_id;
  bool active = chain < a.n_chains;
  const int lane = (int)(tix & 31u);
  const int chains_in_warp = 32 / lpc;
  const int chain_in_warp = lane / lpc;
  const size_t warp_elems = (size_t)chains_in_warp * a.d_pad;
  T* warp_pos = smem + (size_t)(tix >> 5) * 2 * warp_elems;
  T* pos_row = warp_pos + (size_t)chain_in_warp * a.d_pad;
  T* row = warp_pos + warp_elems + (size_t)chain_in_warp * a.d_pad;
  unsigned long long gchain = a.chain_offset + chain;
  const size_t d = (size_t)a.d;
  const unsigned wd = (unsigned)lpc * EPLP;   // lane-padded workspace vector length (32-bit offset arithmetic)
  const size_t cw = slot_id;

  T* e_gm = a.ws_edges + (cw * 6 + 2) * wd;     // edge gradients: only read back on a direction flip
  T* e_gp = a.ws_edges + (cw * 6 + 5) * wd;
  T* w_first = a.ws_first + cw * (size_t)a.cap * 2 * wd;
  T* w_prime = a.ws_prime + cw * (size_t)a.cap * wd;

  size_t smem_off = (size_t)(kHmcBlock >> 5) * 2 * warp_elems;
  smem_off = (smem_off + 3) / 4 * 4;
  // mixture: component means and log weights staged once per CTA in shared memory, lane-padded for vector loads
  TParams<T> tp = a.tp;
  if constexpr (std::is_same<TAG, TagMixture>::value) {
    T* smu = smem + smem_off;
    const int K = a.tp.n_comp;
    const T* gmu = a.tp.dp + K;
    for (int i = (int)tix; i < K * lpc * EPLP; i += blockDim.x) {
      const int k = i / (lpc * EPLP), rem = i - k * (lpc * EPLP);
      const int part = rem / EPLP, j = rem - part * EPLP;
      const int col = part * EPL + j;
      constexpr int VN = VecOf<T>::n;      // unit layout of the lane-padded vectors: unit (j / VN) of lane `part` at (j / VN) * lpc + part
      smu[(size_t)k * lpc * EPLP + ((j / VN) * lpc + part) * VN + (j % VN)] = (j < EPL && col < a.d) ? gmu[(size_t)k * a.d + col] : T(0);
    }
    T* slw = smu + (size_t)K * lpc * EPLP;
    if ((int)tix < K) slw[tix] = a.tp.dp[K + (size_t)K * a.d + tix];
    tp.smem_mu = smu;
    tp.smem_logw = slw;
    smem_off += ((size_t)K * lpc * EPLP + kMaxComp + 3) / 4 * 4;
  }
  T* hot = smem + smem_off + (size_t)(tix / (unsigned)lpc) * H_COUNT * wd;
  // uniform ring of this chain: 2 lpc doubles (kHmcBlock * 2 doubles per CTA), after the hot vectors
  double* s_unif = reinterpret_cast<double*>(smem + smem_off + (size_t)kHmcBlock * H_COUNT * EPLP) + (size_t)(tix / (unsigned)lpc) * 2 * lpc;
  __syncthreads();

  T eps = T(1), eps_bar = T(1), h_bar = T(0), mu = T(0);
  unsigned long long i_norm = 0, i_exp = 0, i_unif = 0;
  const bool inject = a.inj_normals != nullptr;
  unsigned long long my_leapfrogs = 0, my_diverge = 0, my_depth = 0, chain_leaps = 0;
  unsigned int my_moved = 0;
  bool moved = false;      // the current transition accepted at least one subtree proposal
  // Diagonal mass matrix (MASS): its entries are re-read from global memory (L1-resident, coalesced) where they are
  // used instead of living in 2 x EPL registers.
  constexpr bool has_mass = (MASS == 1);
  constexpr bool dense_mass = (MASS == 2);
  const bool dense_on = dense_mass && a.dense_active != 0;
  // dense matrix-vector product for this chain: out_i = sum_j mat[i][j] in_j over j = 0 .. d-1 (or j <= i), summed left to
  // right as MassMatrix::inv_mul / sample_momentum do (generic_nuts.rs:265-303); the vector is exchanged through the
  // chain's scratch row.  Called warp-uniformly.
  auto dense_mul = [&](const T* mat, const T (&in)[EPL], T (&out)[EPL], bool lower) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < EPL; ++j)
      if (j < ln.nvalid) row[ln.lo + j] = in[j];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      T acc = T(0);
      if (e < ln.nvalid) {
        const int i = ln.lo + e;
        const T* mrow = mat + ((size_t)chain * d + (size_t)i) * d;
        const int jmax = lower ? i : a.d - 1;
        for (int jj = 0; jj <= jmax; ++jj) acc = acc + mrow[jj] * row[jj];
      }
      out[e] = acc;
    }
    __syncwarp();
  };
  T mp[EPL];     // M^-1 p of the current leaf (dense mass): shared by the kinetic energy and the whole-trajectory U-turn test
#pragma unroll
  for (int j = 0; j < EPL; ++j) mp[j] = T(0);

  // per-chain state in / out (each lane moves its own slice of the position row)
  auto load_chain = [&]() {
#pragma unroll
    for (int j = 0; j < EPL; ++j)
      if (j < ln.nvalid) pos_row[ln.lo + j] = a.positions[chain * d + ln.lo + j];
    eps = a.eps[chain]; eps_bar = a.eps_bar[chain]; h_bar = a.h_bar[chain]; mu = a.mu[chain];
    if (__builtin_expect(inject, 0)) { i_norm = a.inj_used[chain * 3]; i_exp = a.inj_used[chain * 3 + 1]; i_unif = a.inj_used[chain * 3 + 2]; }
    if (a.write_init && a.out) {
#pragma unroll
      for (int j = 0; j < EPL; ++j)
        if (j < ln.nvalid) a.out[(chain * a.out_n) * d + ln.lo + j] = pos_row[ln.lo + j];
    }
    chain_leaps = 0;
    gchain = a.chain_offset + chain;
  };
  auto store_chain = [&]() {
#pragma unroll
    for (int j = 0; j < EPL; ++j)
      if (j < ln.nvalid) a.positions[chain * d + ln.lo + j] = pos_row[ln.lo + j];
    if (ln.part == 0) {
      a.eps[chain] = eps; a.eps_bar[chain] = eps_bar; a.h_bar[chain] = h_bar;
      if (__builtin_expect(inject, 0)) { a.inj_used[chain * 3] = i_norm; a.inj_used[chain * 3 + 1] = i_exp; a.inj_used[chain * 3 + 2] = i_unif; }
      if (a.chain_leapfrogs) a.chain_leapfrogs[chain] += (long long)chain_leaps;
    }
  };
  if (active) load_chain();
  else {
#pragma unroll
    for (int j = 0; j < EPL; ++j)
      if (j < ln.nvalid) pos_row[ln.lo + j] = pad_q;     // idle lane groups evaluate a harmless point
  }
  __syncwarp();
  bool exhausted = !active;      // no more chains for this lane group

  T q[EPL], p[EPL], g[EPL], prime[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) { q[j] = pad_q; p[j] = T(0); g[j] = T(0); prime[j] = pad_q; }
  {
    // the hot vectors are read unconditionally by the merge loop (lanes that do not merge ignore the result): defined values
    T zero[EPL];
#pragma unroll
    for (int j = 0; j < EPL; ++j) zero[j] = T(0);
#pragma unroll
    for (int h = 0; h < H_COUNT; ++h) store_slice<T, EPL>(hot + (unsigned)h * wd, zero, ln, true);
  }
  int n_stack[kNutsDepthCap];
#pragma unroll
  for (int k = 0; k < kNutsDepthCap; ++k) n_stack[k] = 0;

  int phase = (active && a.n_steps > 0) ? NP_START : NP_DONE;
  if (active && a.n_steps == 0) store_chain();
  uint32_t s = 0;          // transition of this launch
  uint32_t draw = 0;       // uniform draws of the current transition (Philox stream 2)
  int j_depth = 0, v = 1;
  unsigned int leaf_i = 0; // leaf index within the current subtree build
  long long n_tot = 1;
  int nR = 0, n_alpha = 0;
  bool sR = true;
  T alpha_sum = T(0), logu = T(0), joint0 = T(0);
  // Tree uniforms (Philox stream 2: draw i = words (x, y) [i even] or (z, w) [i odd] of block i >> 1).  The lanes of a
  // chain generate lpc blocks = 2 lpc draws at once — Philox costs the same ~70 warp-instructions for one block as for
  // one block per lane — into a shared-memory ring; the draws in between are one shared-memory read.
  const unsigned chain_mask = (lpc >= 32) ? kFull : (((1u << lpc) - 1u) << ln.gbase);
  const uint32_t u_per = 2u * (uint32_t)lpc;
  auto next_unif = [&]() -> double {
    if (__builtin_expect(inject, 0)) {
      const double u = (i_unif < a.n_unif) ? a.inj_unif[chain * a.n_unif + i_unif] : 0.75;
      ++i_unif;
      return u;
    }
    const uint32_t idx = draw & (u_per - 1u);
    if (idx == 0u) {
      __syncwarp(chain_mask);
      const uint4 r = philox4x32_10(philox_ctr(gchain, a.step_base + s, 2u, (draw >> 1) + (uint32_t)ln.part), a.key);
      s_unif[2 * ln.part] = u01d(r.x, r.y);
      s_unif[2 * ln.part + 1] = u01d(r.z, r.w);
      __syncwarp(chain_mask);
    }
    ++draw;
    return s_unif[idx];
  };

  for (;;) {
    // ---- 0. lane groups whose chain is finished take the next one from the queue
    {
      const bool need = (phase == NP_DONE) && !exhausted;
      if (__any_sync(kFull, need)) {
        unsigned long long nxt = 0;
        if (need && ln.part == 0) nxt = atomicAdd(a.queue, 1ull);
        nxt = __shfl_sync(kFull, nxt, ln.gbase);
        if (need) {
          if (nxt < (unsigned long long)a.n_chains) {
            chain = (size_t)nxt;
            load_chain();
            s = 0;
            if (a.n_steps > 0) phase = NP_START; else store_chain();
          } else {
            exhausted = true;
          }
        }
        __syncwarp();
      }
      if (__all_sync(kFull, phase == NP_DONE && exhausted)) break;
    }
    // ---- A. momentum for the chains that start a transition (generic_nuts.rs:759-762)
    const bool is_start = (phase == NP_START);
    const bool is_leaf = (phase == NP_LEAF);
    if (__any_sync(kFull, is_start)) {
      T pn[EPL];
      if (__builtin_expect(inject, 0)) {
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
          const unsigned long long idx = i_norm + (unsigned long long)(ln.lo + j);
          pn[j] = (is_start && j < ln.nvalid) ? (T)((idx < a.n_norm) ? a.inj_normals[chain * a.n_norm + idx] : 0.0) : T(0);
        }
        if (is_start) i_norm += d;
      } else {
        constexpr int NPB = NormalsPerBlock<T>::value;
        const int nblocks = (a.d + NPB - 1) / NPB;
        for (int b = ln.part; b < nblocks; b += lpc) {
          T z[NPB];
          normals_from_block<kExact>(philox4x32_10(philox_ctr(gchain, a.step_base + s, 0u, (uint32_t)b), a.key), z);
#pragma unroll
          for (int k = 0; k < NPB; ++k)
            if (b * NPB + k < a.d_pad) row[b * NPB + k] = z[k];
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < EPL; ++j) pn[j] = (j < ln.nvalid) ? row[ln.lo + j] : T(0);
        __syncwarp();
      }
      if (is_start) {   // sample_momentum, generic_nuts.rs:283-303: z * sqrt(var)
#pragma unroll
        for (int j = 0; j < EPL; ++j) { p[j] = pn[j]; q[j] = (j < ln.nvalid) ? pos_row[ln.lo + j] : pad_q; }
        if constexpr (has_mass) {
#pragma unroll
          for (int j = 0; j < EPL; ++j)
            if (j < ln.nvalid) p[j] = pn[j] * a.mass_sqrt[chain * d + ln.lo + j];
        }
      }
      if constexpr (dense_mass) {
        if (dense_on) {      // sample_momentum, Dense: p = chol z (generic_nuts.rs:292-301)
          T pc[EPL];
          dense_mul(a.mass_chol, pn, pc, true);
          if (is_start) {
#pragma unroll
            for (int j = 0; j < EPL; ++j) p[j] = pc[j];
          }
        }
      }
    }

    // ---- B. TWO gradient evaluations per chain and pass.  A chain that starts a transition evaluates the gradient at its
    // position (slot 0), then takes the single leaf of the first doubling (slot 1); afterwards every doubling has an even
    // number of leaves and a pass is one pair: leaf A (even index, slot 0: it becomes the pending left subtree of level 0) and
    // leaf B (odd, slot 1: merged with A below).  The chains of a warp therefore stay aligned on pair boundaries, the even
    // leaf never enters the merge loop, and the per-pass control (queue, momentum, phase tests) is paid once per two leaves.
    bool in_merge = false;
    bool a_ok = true;        // leaf A kept its subtree alive (s'): leaf B is built (generic_nuts.rs:1251)
#pragma unroll 1
    for (int slot = 0; slot < 2; ++slot) {
      const bool mv = (slot == 0) ? is_leaf : (is_start || (is_leaf && a_ok));   // this chain takes a leapfrog in this slot
      const T veps = (T)v * eps;                 // generic_nuts.rs:1187
      const T he = veps * T(0.5);                // leapfrog_with_mass :1396-1418
      if (mv) {
#pragma unroll
        for (int j = 0; j < EPL; ++j) p[j] = p[j] + g[j] * he;
      }
      if constexpr (dense_mass) {
        if (dense_on) dense_mul(a.mass_dinv, p, mp, false);                      // velocity = M^-1 p (apply_inv_mass)
      }
      if (mv) {
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
          if constexpr (has_mass) {                                             // velocity = M^-1 p (apply_inv_mass)
            const T mi = (j < ln.nvalid) ? a.mass_inv[chain * d + ln.lo + j] : T(1);
            q[j] = q[j] + (mi * p[j]) * veps;
          } else if constexpr (dense_mass) {
            q[j] = q[j] + (dense_on ? mp[j] : p[j]) * veps;
          } else {
            q[j] = q[j] + p[j] * veps;
          }
        }
      }
      // the gradient is written in place: a chain that does not move in this slot re-evaluates it at the same point (same
      // value), a finished chain never reads it again
      const T logp = eval_target<T, EPL, EP, true>(TAG{}, q, g, ln, tp, row);
      if (mv) {
#pragma unroll
        for (int j = 0; j < EPL; ++j) p[j] = p[j] + g[j] * he;
      }
      if constexpr (dense_mass) {
        if (dense_on) dense_mul(a.mass_dinv, p, mp, false);                      // row_dot of MassMatrix::kinetic, Dense
      }
      T terms[EPL];
#pragma unroll
      for (int j = 0; j < EPL; ++j) {                                              // MassMatrix::kinetic :228-263
        if constexpr (has_mass) terms[j] = p[j] * p[j] * ((j < ln.nvalid) ? a.mass_inv[chain * d + ln.lo + j] : T(1));
        else if constexpr (dense_mass) terms[j] = dense_on ? p[j] * mp[j] : p[j] * p[j];
        else terms[j] = p[j] * p[j];
      }
      const T ke = T(0.5) * chain_sum<T, EPL>(terms, ln);
      const T joint = logp - ke;

      if (slot == 0 && is_start) {
        // generic_nuts.rs:765-781
        joint0 = joint;
        T e1;
        if (__builtin_expect(inject, 0)) { e1 = (T)((i_exp < a.n_exp) ? a.inj_exp1[chain * a.n_exp + i_exp] : 1.0); ++i_exp; }
        else {
          const uint4 r = philox4x32_10(philox_ctr(gchain, a.step_base + s, 1u, 0u), a.key);
          if constexpr (!kExact && sizeof(T) == 4) e1 = -__logf(u01(r.z));   // fast mode, f32: the 24 leading bits of word 2
          else e1 = (T)(-log(u01d(r.z, r.w)));
        }
        logu = joint0 - e1;
        store_slice<T, EPL>(hot + H_QM * wd, q, ln, true); store_slice<T, EPL>(hot + H_PM * wd, p, ln, true);
        store_slice<T, EPL>(hot + H_QP * wd, q, ln, true); store_slice<T, EPL>(hot + H_PP * wd, p, ln, true);
        store_slice<T, EPL>(e_gm, g, ln, true); store_slice<T, EPL>(e_gp, g, ln, true);
        j_depth = 0; n_tot = 1; draw = 0; moved = false;
        const T u1 = (T)next_unif();             // :783-784
        v = (u1 < T(0.5)) ? 1 : -1;
        leaf_i = 0; alpha_sum = T(0); n_alpha = 0;
        phase = NP_LEAF;
      } else if (mv) {
        // leaf of build_tree (j == 0 branch, generic_nuts.rs:1185-1222)
        ++my_leapfrogs; ++chain_leaps;
        nR = (logu < joint) ? 1 : 0;
        sR = (logu - T(1000)) < joint;
        if (!sR) ++my_diverge;
        alpha_sum = alpha_sum + min(T(1), fast_exp<T>(joint - joint0));
        ++n_alpha;
#pragma unroll
        for (int j = 0; j < EPL; ++j) prime[j] = q[j];
        if (slot == 0) {
          // leaf A: the first leaf of the left subtree of the level-0 merge that follows in this pass (shared memory) and,
          // when its index is a multiple of 4, of later merges at levels >= 1 (global workspace, stack slot popcount(c >> 1))
          store_slice<T, EPL>(hot + H_LQ * wd, q, ln, true);
          store_slice<T, EPL>(hot + H_LP * wd, p, ln, true);
          if ((leaf_i & 3u) == 0u) {
            const int slot_w = __popc(leaf_i >> 1);
            store_slice<T, EPL>(w_first + (unsigned)(slot_w * 2) * wd, q, ln, true);
            store_slice<T, EPL>(w_first + (unsigned)(slot_w * 2 + 1) * wd, p, ln, true);
          }
          a_ok = sR;
          if (sR) { n_stack[0] = nR; ++leaf_i; }    // pending left subtree of level 0; leaf B follows
          else in_merge = true;                     // failed subtree: passed up through the left children to the top
        } else {
          in_merge = true;                          // leaf B (level-0 merge first) or the single leaf of doubling 0 (top)
        }
      }
    }
    __syncwarp();     // the slices written above are read back (by the same lanes) through a different pointer type below

    // ---- C. merges with the pending left subtrees, then the top-level step of the doubling
    int k = 0;
    while (__any_sync(kFull, in_merge)) {
      bool do_merge = false, do_top = false;
      if (in_merge) {
        if (k == j_depth) do_top = true;                       // the depth-j subtree is complete (or failed)
        else if ((leaf_i >> k) & 1u) do_merge = true;          // pending left sibling at level k
        else if (sR) {                                         // becomes the pending left subtree of level k
          n_stack[k] = nR;
          if (k >= 1) store_slice<T, EPL>(w_prime + (unsigned)k * wd, prime, ln, true);
          in_merge = false;
          ++leaf_i;
        }                                                      // else: failed subtree passed up through a left child
      }
      // the other end of the U-turn test: first leaf of the left subtree (merge) / the other trajectory edge (top).
      // Lanes that do neither read the (always defined) last-even-leaf vectors and ignore the result.
      T fq[EPL], fp[EPL];
      const T* src_q = hot + H_LQ * wd;
      const T* src_p = hot + H_LP * wd;
      if (do_merge && k > 0) {
        const unsigned int start = (leaf_i >> (k + 1)) << (k + 1);
        const int slot = __popc(start >> 1);
        src_q = w_first + (unsigned)(slot * 2) * wd;
        src_p = w_first + (unsigned)(slot * 2 + 1) * wd;
      } else if (do_top) {
        src_q = hot + (unsigned)((v == 1) ? H_QM : H_QP) * wd;
        src_p = hot + (unsigned)((v == 1) ? H_PM : H_PP) * wd;
      }
      load_slice_raw<T, EPL>(fq, src_q, ln);
      load_slice_raw<T, EPL>(fp, src_p, ln);
      // stop_criterion (generic_nuts.rs:1357-1378, identity mass): diff = q+ - q- ; diff.p- >= 0 && diff.p+ >= 0
      // With diff = q - f (current leaf minus the loaded end), r1 = diff . v(f), r2 = diff . v(current):
      //   forward  (v = +1: f is the minus end): (q+ - q-) . v- = r1, (q+ - q-) . v+ = r2  -> both >= 0
      //   backward (v = -1: f is the plus end):  (q+ - q-) = -diff: . v- = -r2, . v+ = -r1  -> both r <= 0
      // (negation is exact and commutes with rounding, so the sums are bit-identical to the reference's order)
      const bool fwd = (v == 1);
      T vf[EPL];      // M^-1 p of the other trajectory end (dense mass, whole-trajectory test only: sub-tree tests use the identity, :1316)
#pragma unroll
      for (int j = 0; j < EPL; ++j) vf[j] = T(0);
      if constexpr (dense_mass) {
        if (dense_on && __any_sync(kFull, do_top)) dense_mul(a.mass_dinv, fp, vf, false);
      }
      T r1, r2;
      chain_sum_fn2<T, EPL>(ln, [&](int j) {
        const T df = q[j] - fq[j];
        if (has_mass && do_top && j < ln.nvalid) return df * (a.mass_inv[chain * d + ln.lo + j] * fp[j]);
        if (dense_mass && dense_on && do_top) return (j < ln.nvalid) ? df * vf[j] : T(0);
        return (!EP || j < ln.nvalid) ? df * fp[j] : T(0);
      }, [&](int j) {
        const T df = q[j] - fq[j];
        if (has_mass && do_top && j < ln.nvalid) return df * (a.mass_inv[chain * d + ln.lo + j] * p[j]);
        if (dense_mass && dense_on && do_top) return (j < ln.nvalid) ? df * mp[j] : T(0);
        return (!EP || j < ln.nvalid) ? df * p[j] : T(0);
      }, r1, r2);
      const bool crit = fwd ? ((r1 >= T(0)) && (r2 >= T(0))) : ((r1 <= T(0)) && (r2 <= T(0)));
      if (do_merge) {
        // generic_nuts.rs:1305-1323
        const double u = next_unif();
        const int nL = n_stack[k];
        const int den = (nL + nR) > 1 ? (nL + nR) : 1;
        bool take_right;
        if constexpr (kExact) take_right = u < ((double)nR / (double)den);
        else take_right = u * (double)den < (double)nR;       // same test without the f64 division
        if (!take_right) {
          if (k == 0) {
#pragma unroll
            for (int j = 0; j < EPL; ++j) prime[j] = fq[j];
          } else {
            load_slice_raw<T, EPL>(prime, w_prime + (unsigned)k * wd, ln);
          }
        }
        nR += nL;
        sR = sR && crit;
      } else if (do_top) {
        // generic_nuts.rs:803-880: new edge, accept the subtree's proposal, trajectory-level U-turn test
        store_slice<T, EPL>(hot + (unsigned)((v == 1) ? H_QP : H_QM) * wd, q, ln, true);
        store_slice<T, EPL>(hot + (unsigned)((v == 1) ? H_PP : H_PM) * wd, p, ln, true);
        store_slice<T, EPL>((v == 1) ? e_gp : e_gm, g, ln, true);
        const T ratio = (T)nR / (T)n_tot;
        const T tmp = ratio < T(1) ? ratio : T(1);
        const T u2 = (T)next_unif();
        if (sR && (u2 < tmp)) {
#pragma unroll
          for (int j = 0; j < EPL; ++j)
            if (j < ln.nvalid) pos_row[ln.lo + j] = prime[j];
          moved = true;
        }
        n_tot += nR;
        bool cont = sR && crit;
        ++j_depth;
        if (j_depth >= a.max_depth) cont = false;
        in_merge = false;
        if (cont) {
          const T u1 = (T)next_unif();
          const int vn = (u1 < T(0.5)) ? 1 : -1;
          if (vn != v) {
            load_slice_raw<T, EPL>(q, hot + (unsigned)((vn == 1) ? H_QP : H_QM) * wd, ln);
            load_slice_raw<T, EPL>(p, hot + (unsigned)((vn == 1) ? H_PP : H_PM) * wd, ln);
            load_slice_raw<T, EPL>(g, vn == 1 ? e_gp : e_gm, ln);
          }
          v = vn;
          leaf_i = 0; alpha_sum = T(0); n_alpha = 0;
        } else {
          phase = NP_END;
        }
      }
      ++k;
    }

    // ---- D. end of the transition: dual averaging (generic_nuts.rs:882-924), write-out
    bool did_collect = false;     // dense mass: this chain added its position to the running covariance in this pass
    T dlt[dense_mass ? EPL : 1];
    if (phase == NP_END) {
      const uint32_t m = a.m_base + s + 1;
      my_depth += (unsigned long long)j_depth;
      my_moved += moved ? 1u : 0u;
      T eta = T(1) / (T)(m + 10u);
      h_bar = (T(1) - eta) * h_bar + eta * (a.target_accept - alpha_sum / (T)n_alpha);
      if (m <= a.n_discard) {
        const T mm = (T)m;
        eps = exp(mu - sqrt(mm) / T(0.05) * h_bar);
        if constexpr (!kExact && sizeof(T) == 4) eta = rsqrtf(mm) * rsqrtf(sqrtf(mm));   // m^(-3/4)
        else eta = pow(mm, -T(0.75));
        eps_bar = exp((T(1) - eta) * log(eps_bar) + eta * log(eps));
      } else {
        eps = eps_bar;
      }
      if (MASS && a.run_mean && m <= a.n_discard && m > a.collect_after && m < a.collect_before) {
        // RunningCov::update (generic_nuts.rs:105-114) on the position after this transition
        // every chain collects at the same transitions, so the count is a function of m alone (no device counter)
        const uint32_t first_m = a.m_base > a.collect_after ? a.m_base : a.collect_after;
        const T n_s = (T)(a.run_n_base + (m - first_m));
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
          if (j < ln.nvalid) {
            const size_t idx = chain * d + ln.lo + j;
            const T x = pos_row[ln.lo + j];
            T mean = a.run_mean[idx];
            const T delta = x - mean;
            mean = mean + delta / n_s;
            const T delta2 = x - mean;
            a.run_mean[idx] = mean;
            a.run_m2[idx] = a.run_m2[idx] + delta * delta2;
            if constexpr (dense_mass) { dlt[j] = delta; row[ln.lo + j] = delta2; }
          }
        }
        did_collect = true;
      }
      const long long slot = (long long)m - a.rec_off;
      if (a.out && slot >= 0 && slot < (long long)a.out_n) {
#pragma unroll
        for (int j = 0; j < EPL; ++j)
          if (j < ln.nvalid) __stcs(a.out + (chain * a.out_n + (size_t)slot) * d + ln.lo + j, pos_row[ln.lo + j]);
      }
      ++s;
      if (s < a.n_steps) phase = NP_START;
      else { phase = NP_DONE; store_chain(); }
    }
    if constexpr (dense_mass) {
      // RunningCov::update, dense part (generic_nuts.rs:115-126): m2[i][j] += delta_i * delta2_j for j >= i; the chain's
      // delta2 vector was left in its scratch row above
      if (a.run_m2d && __any_sync(kFull, did_collect)) {
        __syncwarp();
        if (did_collect) {
#pragma unroll
          for (int e = 0; e < EPL; ++e) {
            if (e < ln.nvalid) {
              const int i = ln.lo + e;
              T* mrow = a.run_m2d + ((size_t)chain * d + (size_t)i) * d;
              for (int jj = i; jj < a.d; ++jj) mrow[jj] = mrow[jj] + dlt[e] * row[jj];
            }
          }
        }
        __syncwarp();
      }
    }
  }

  if (ln.part != 0) { my_leapfrogs = 0; my_diverge = 0; my_depth = 0; my_moved = 0; }
  for (int o = 16; o > 0; o >>= 1) {
    my_leapfrogs += __shfl_xor_sync(kFull, my_leapfrogs, o);
    my_diverge += __shfl_xor_sync(kFull, my_diverge, o);
    my_depth += __shfl_xor_sync(kFull, my_depth, o);
    my_moved += __shfl_xor_sync(kFull, my_moved, o);
  }
  if (lane == 0) {
    if (my_moved && a.accept_total) atomicAdd(a.accept_total, (unsigned long long)my_moved);
    if (my_leapfrogs) atomicAdd(a.leapfrog_total, my_leapfrogs);
    if (my_diverge) atomicAdd(a.diverge_total, my_diverge);
    if (my_depth) atomicAdd(a.depth_total, my_depth);
  }
}

// ----------------------------------------------------------------------------------------------
// init_chain_state (generic_nuts.rs:731-753): consume d normals; find_reasonable_epsilon when the step
// size is still the -1 sentinel; mu = ln(10 eps).
// ----------------------------------------------------------------------------------------------
template <class T>
struct NutsInitArgs {
  TParams<T> tp;
  size_t n_chains;
  unsigned long long chain_offset;
  PhiloxKey key;
  uint32_t step;          // Philox transition index reserved for this init
  const T* positions;
  int d, d_pad, lpc;
  T* eps; T* mu;
  const double* inj_normals; size_t n_norm;
  unsigned long long* inj_used;
  // diagonal mass matrix: the momentum is z * sqrt(var) (sample_momentum, generic_nuts.rs:283-303); the step-size
  // search itself runs with the identity mass, as the reference's find_reasonable_epsilon does (:1009-1023)
  const T* mass_sqrt;
  const T* mass_chol;     // dense mass: lower Cholesky factor [C, d, d], momentum = chol z (or null)
  // probe = 1: the step-size reset after a mass-matrix update (generic_nuts.rs:906-918): always search, from a fresh
  // momentum (Philox stream 3 of transition `step`), then mu = ln(10 eps), eps_bar = eps, h_bar = 0
  int probe;
  T* eps_bar; T* h_bar;
};

template <class T, int EPL, class TAG>
__global__ void __launch_bounds__(kHmcBlock, 2) nuts_init_kernel(const NutsInitArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const Lane ln = make_lane<EPL>(tid, a.lpc, a.d);
  const size_t chain = (size_t)(tid / a.lpc);
  const bool active = chain < a.n_chains;
  const int lane = threadIdx.x & 31;
  const int chain_in_warp = lane / a.lpc;
  T* row = smem + ((size_t)(threadIdx.x >> 5) * (32 / a.lpc) + chain_in_warp) * a.d_pad;
  const unsigned long long gchain = a.chain_offset + chain;
  const size_t d = (size_t)a.d;

  T pos[EPL], mom[EPL], g0[EPL], q1[EPL], p1[EPL], g1[EPL];
#pragma unroll
  for (int j = 0; j < EPL; ++j) pos[j] = (active && j < ln.nvalid) ? a.positions[chain * d + ln.lo + j] : T(1);
  if (a.inj_normals) {
    unsigned long long i_norm = active ? a.inj_used[chain * 3] : 0;
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      const unsigned long long idx = i_norm + (unsigned long long)(ln.lo + j);
      mom[j] = (active && j < ln.nvalid) ? (T)((idx < a.n_norm) ? a.inj_normals[chain * a.n_norm + idx] : 0.0) : T(0);
    }
    if (active && ln.part == 0) a.inj_used[chain * 3] = i_norm + d;
  } else {
    constexpr int NPB = NormalsPerBlock<T>::value;
    const int nblocks = (a.d + NPB - 1) / NPB;
    for (int b = ln.part; b < nblocks; b += a.lpc) {
      T z[NPB];
      normals_from_block<kExact>(philox4x32_10(philox_ctr(gchain, a.step, a.probe ? 3u : 0u, (uint32_t)b), a.key), z);
#pragma unroll
      for (int k = 0; k < NPB; ++k)
        if (b * NPB + k < a.d_pad) row[b * NPB + k] = z[k];
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < EPL; ++j) mom[j] = (j < ln.nvalid) ? row[ln.lo + j] : T(0);
    __syncwarp();
  }
  if (a.mass_sqrt) {
#pragma unroll
    for (int j = 0; j < EPL; ++j)
      if (active && j < ln.nvalid) mom[j] = mom[j] * a.mass_sqrt[chain * d + ln.lo + j];
  }
  if (a.mass_chol) {      // sample_momentum, Dense (generic_nuts.rs:292-301): chol z, summed left to right
    __syncwarp();
#pragma unroll
    for (int j = 0; j < EPL; ++j)
      if (j < ln.nvalid) row[ln.lo + j] = mom[j];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      T acc = T(0);
      if (active && e < ln.nvalid) {
        const int i = ln.lo + e;
        const T* mrow = a.mass_chol + (chain * d + (size_t)i) * d;
        for (int jj = 0; jj <= i; ++jj) acc = acc + mrow[jj] * row[jj];
      }
      mom[e] = acc;
    }
    __syncwarp();
  }
  T eps = active ? a.eps[chain] : T(1);
  const bool need = active && (a.probe || fabs(eps + T(1)) <= (sizeof(T) == 4 ? T(1.1920929e-07) : T(2.220446049250313e-16)));

  // find_reasonable_epsilon_with_mass, generic_nuts.rs:1025-1102 (identity mass).  stage 0: gradient at
  // the position; stage 1: halve until finite; stage 2: double / halve until the acceptance crosses 1/2.
  int stage = need ? 0 : 3;
  T epsilon = T(1), kfac = T(1), ulogp = T(0), ke_mom = T(0), afac = T(1), e_try = T(1);
  const T half = T(0.5);
  const T ln_half = (sizeof(T) == 4) ? (T)-0.6931472f : (T)-0.6931471805599453;
  const T ln_two = (sizeof(T) == 4) ? (T)0.6931472f : (T)0.6931471805599453;
  int guard = 0;
  while (__any_sync(kFull, stage < 3)) {
    if (stage == 0) {
#pragma unroll
      for (int j = 0; j < EPL; ++j) { q1[j] = pos[j]; p1[j] = mom[j]; }
    } else {
      const T he = e_try * half;
#pragma unroll
      for (int j = 0; j < EPL; ++j) { p1[j] = mom[j] + g0[j] * he; q1[j] = pos[j] + p1[j] * e_try; }
    }
    const T lp = eval_target<T, EPL, true, true>(TAG{}, q1, g1, ln, a.tp, row);
    T terms[EPL], fin[EPL];
    if (stage != 0) {
      const T he = e_try * half;
#pragma unroll
      for (int j = 0; j < EPL; ++j) p1[j] = p1[j] + g1[j] * he;
    }
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      terms[j] = p1[j] * p1[j];
      fin[j] = (j < ln.nvalid && !(fabs(g1[j]) < T(INFINITY))) ? T(1) : T(0);   // counts NaN and +-inf
    }
    const T ke = half * chain_sum<T, EPL>(terms, ln);
    const T bad = chain_sum<T, EPL>(fin, ln);
    const bool finite = (fabs(lp) < T(INFINITY)) && (bad == T(0));
    if (stage == 0) {
      ulogp = lp; ke_mom = ke;
#pragma unroll
      for (int j = 0; j < EPL; ++j) g0[j] = g1[j];
      e_try = epsilon;            // first trial: leapfrog with epsilon = 1
      stage = 1;
    } else if (stage == 1) {
      if (!finite && guard < 200) { kfac = kfac * half; e_try = epsilon * kfac; ++guard; }
      else {
        epsilon = half * kfac * epsilon;
        const T lap = lp - ulogp - (ke - ke_mom);
        afac = (lap > ln_half) ? T(1) : T(-1);
        if (afac * lap > -afac * ln_two) { epsilon = epsilon * (afac > T(0) ? T(2) : half); e_try = epsilon; stage = 2; guard = 0; }
        else stage = 3;
      }
    } else if (stage == 2) {
      const T lap = lp - ulogp - (ke - ke_mom);
      if (afac * lap > -afac * ln_two && guard < 200) { epsilon = epsilon * (afac > T(0) ? T(2) : half); e_try = epsilon; ++guard; }
      else stage = 3;
    }
  }
  if (active && ln.part == 0) {
    if (need) { eps = epsilon; a.eps[chain] = eps; }
    a.mu[chain] = log(T(10) * eps);
    if (a.probe) { a.eps_bar[chain] = eps; a.h_bar[chain] = T(0); }
  }
}

template <class T>
inline NutsArgs<T> make_nuts_args(const NutsLaunch& L) {
  NutsArgs<T> a;
  a.tp = make_tparams<T>(L.tgt);
  a.n_chains = L.n_chains; a.chain_offset = L.chain_offset;
  a.key = PhiloxKey{(uint32_t)L.seed, (uint32_t)(L.seed >> 32)};
  a.step_base = L.step_base;
  a.positions = (T*)L.positions;
  a.d = L.tgt.dim; a.d_pad = ((L.tgt.dim + 3) / 4) * 4; a.lpc = L.lpc;
  a.n_steps = L.n_steps; a.m_base = L.m_base; a.n_discard = L.n_discard;
  a.rec_off = L.rec_off; a.write_init = L.write_init;
  a.out = (T*)L.out; a.out_n = L.out_n;
  a.eps = (T*)L.eps; a.eps_bar = (T*)L.eps_bar; a.h_bar = (T*)L.h_bar; a.mu = (T*)L.mu;
  a.target_accept = (T)L.target_accept;
  a.max_depth = L.max_depth;
  a.ws_edges = (T*)L.ws_edges; a.ws_first = (T*)L.ws_first; a.ws_prime = (T*)L.ws_prime; a.cap = L.cap;
  a.leapfrog_total = L.leapfrog_total; a.diverge_total = L.diverge_total; a.depth_total = L.depth_total;
  a.accept_total = L.accept_total;
  a.chain_leapfrogs = L.chain_leapfrogs;
  a.inj_normals = L.inj_normals; a.n_norm = L.n_norm; a.inj_exp1 = L.inj_exp1; a.n_exp = L.n_exp;
  a.inj_unif = L.inj_unif; a.n_unif = L.n_unif; a.inj_used = L.inj_used;
  a.queue = L.queue;
  a.mass_inv = (const T*)L.mass_inv; a.mass_sqrt = (const T*)L.mass_sqrt;
  a.run_mean = (T*)L.run_mean; a.run_m2 = (T*)L.run_m2; a.run_n_base = L.run_n_base;
  a.collect_after = L.collect_after; a.collect_before = L.collect_before;
  a.mass_dinv = (const T*)L.mass_dinv; a.mass_chol = (const T*)L.mass_chol; a.run_m2d = (T*)L.run_m2d;
  a.dense_active = L.dense_active;
  return a;
}

template <class T, int EPL, class TAG>
inline cudaError_t nuts_launch_one(const NutsLaunch& L, cudaStream_t st) {
  const int d_pad = ((L.tgt.dim + 3) / 4) * 4;
  const size_t threads = L.n_chains * (size_t)L.lpc;
  const unsigned blocks = (unsigned)((threads + kHmcBlock - 1) / kHmcBlock);
  if (L.init_only) {
    NutsInitArgs<T> a;
    a.tp = make_tparams<T>(L.tgt);
    a.n_chains = L.n_chains; a.chain_offset = L.chain_offset;
    a.key = PhiloxKey{(uint32_t)L.seed, (uint32_t)(L.seed >> 32)};
    a.step = L.step_base; a.positions = (const T*)L.positions;
    a.d = L.tgt.dim; a.d_pad = d_pad; a.lpc = L.lpc;
    a.eps = (T*)L.eps; a.mu = (T*)L.mu;
    a.inj_normals = L.inj_normals; a.n_norm = L.n_norm; a.inj_used = L.inj_used;
    a.mass_sqrt = (const T*)L.mass_sqrt; a.probe = L.probe; a.eps_bar = (T*)L.eps_bar; a.h_bar = (T*)L.h_bar;
    a.mass_chol = (L.mass_chol && L.dense_active) ? (const T*)L.mass_chol : nullptr;
    const size_t smem = (size_t)(kHmcBlock / L.lpc) * d_pad * sizeof(T);
    auto kern = nuts_init_kernel<T, EPL, TAG>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    kern<<<blocks, kHmcBlock, smem, st>>>(a);
    return cudaGetLastError();
  }
  NutsArgs<T> a = make_nuts_args<T>(L);
  // shared memory (elements of T): position + scratch rows, [mixture means + log weights], hot per-chain vectors
  size_t smem_el = ((size_t)(kHmcBlock >> 5) * 2 * (32 / L.lpc) * d_pad + 3) / 4 * 4;
  if (std::is_same<TAG, TagMixture>::value) smem_el += ((size_t)L.tgt.n_comp * L.lpc * Eplp<EPL>::value + kMaxComp + 3) / 4 * 4;
  smem_el += (size_t)kHmcBlock * H_COUNT * Eplp<EPL>::value;
  const size_t smem = smem_el * sizeof(T) + (size_t)kHmcBlock * 2 * sizeof(double);   // + the tree-uniform rings
  const bool exact_fit = (L.epl * L.lpc == L.tgt.dim);
  // lanes per chain as a compile-time constant for the wide 8-coordinate layout (d = 65 .. 128: BASELINE config 5)
  constexpr int kLpcFixed = (EPL == 8) ? 16 : 0;
  const bool fixed = kLpcFixed > 0 && L.lpc == kLpcFixed;
  const int mass = L.mass_dinv ? 2 : (L.mass_inv ? 1 : 0);
  void (*kern)(const NutsArgs<T>) = nullptr;
  if (mass == 2) {
    // dense mass matrices are instantiated for the narrow slices only (d <= 128 covers dense_max_dim = 75 with room to spare)
    if constexpr (EPL <= 8) kern = exact_fit ? nuts_run_kernel<T, EPL, TAG, false, 2, 0> : nuts_run_kernel<T, EPL, TAG, true, 2, 0>;
    else return cudaErrorInvalidValue;
  } else if (fixed) kern = mass ? (exact_fit ? nuts_run_kernel<T, EPL, TAG, false, 1, kLpcFixed> : nuts_run_kernel<T, EPL, TAG, true, 1, kLpcFixed>)
                              : (exact_fit ? nuts_run_kernel<T, EPL, TAG, false, 0, kLpcFixed> : nuts_run_kernel<T, EPL, TAG, true, 0, kLpcFixed>);
  else kern = mass ? (exact_fit ? nuts_run_kernel<T, EPL, TAG, false, 1, 0> : nuts_run_kernel<T, EPL, TAG, true, 1, 0>)
                   : (exact_fit ? nuts_run_kernel<T, EPL, TAG, false, 0, 0> : nuts_run_kernel<T, EPL, TAG, true, 0, 0>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  // persistent grid: as many CTAs as fit on the device at once (or fewer when there are fewer chains); the lane groups
  // pull further chains from the queue, which starts at the number of groups handed out statically
  int per_sm = 0, dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t eo = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kHmcBlock, smem);
  if (eo != cudaSuccess) return eo;
  const unsigned resident = (unsigned)(per_sm > 0 ? per_sm : 1) * (unsigned)sms;
  const unsigned grid = blocks < resident ? blocks : resident;
  const unsigned long long first = (unsigned long long)grid * (kHmcBlock / L.lpc);
  cudaError_t eq = cudaMemcpyAsync(L.queue, &first, sizeof first, cudaMemcpyHostToDevice, st);
  if (eq != cudaSuccess) return eq;
  cudaStreamSynchronize(st);   // `first` is a stack variable
  kern<<<grid, kHmcBlock, smem, st>>>(a);
  return cudaGetLastError();
}

// EPL menu of the NUTS kernels: {2 (2-D targets), 4, 8, 25 (f32 only), 13 (f64 only)}
template <class T, class TAG>
inline cudaError_t nuts_dispatch(const NutsLaunch& L, cudaStream_t st) {
  switch (L.epl) {
    case 4: return nuts_launch_one<T, 4, TAG>(L, st);
    case 8: return nuts_launch_one<T, 8, TAG>(L, st);
    case 13:
      if constexpr (sizeof(T) == 8) return nuts_launch_one<T, 13, TAG>(L, st);
      else return cudaErrorInvalidValue;
    case 25:
      if constexpr (sizeof(T) == 4) return nuts_launch_one<T, 25, TAG>(L, st);
      else return cudaErrorInvalidValue;
  }
  return cudaErrorInvalidValue;
}

}  // namespace GM_NS
}  // namespace gm
