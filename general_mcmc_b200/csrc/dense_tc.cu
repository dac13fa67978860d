// dense_tc.cu — K3: HMC on an N-D dense-covariance Gaussian with the gradient GEMM on the 5th-generation
// tensor cores (tcgen05 + TMEM + TMA, sm_100a).
//
// Reference path replaced: BatchedGenericHMC::step / leapfrog (/root/reference/src/batched_hmc.rs:129-190)
// with the target of DiffableGaussian2D::unnorm_logp_batch generalised to d dimensions
// (/root/reference/src/distributions.rs:265-291):  delta = x - mu ; z = delta . P ; logp = c - 1/2 sum(z * delta) ;
// grad = -z.  z for all chains is a real [chains x d] . [d x d] GEMM per gradient evaluation.
//
// Precision: plain TF32 / FP16 (11-bit significand) cannot meet the 1e-5 per-step bar, so both operands are split
// x = hi + lo (22 bits) and the accumulator receives hi.hi + lo.hi + hi.lo in FP32 (TMEM): three MMAs per K-step.
//   GM_TC_F16 = 1 (default): hi = half(x), lo = half(x - hi), tcgen05.mma.kind::f16 (K = 16 per instruction, twice
//     the TF32 rate).  FP16 products are exact in FP32.  P is pre-scaled by a power of two so that its entries use
//     the upper part of FP16's exponent range (the epilogue undoes it); delta = q - mu is scaled likewise, per
//     transition, by the power of two that puts max |delta| of the batch at 2^8 (dense_begin_kernel reduces the maximum,
//     dense_scale_kernel turns it into the scale the GEMM launches read), so targets of any variance scale (1e-8 ..
//     1e+10) stay inside FP16's range with 2^7 headroom for the trajectory; a low part that falls into the subnormals
//     costs at most 2^-32 of max |delta|.  Measured against the f64 oracle at d = 1000,
//     L = 32: same error as the TF32 split (tests/test_gpu_dense_tc.py holds both to 2e-5); 7.5e7 -> 9.3e7 grad-evals/s
//     (1.1e8 with the persistent unit schedule below).
//   GM_TC_F16 = 0: hi = rna_tf32(x), lo = rna_tf32(x - hi), tcgen05.mma.kind::tf32 (K = 8).
//
// State: the trajectory runs in delta-space.  p [C, d] and delta [C, kpad] (two ping-pong buffers, padded columns
// zero) are the only per-leapfrog HBM arrays: 16 bytes per coordinate per leapfrog.  One transition (fast math
// mode, merged half kicks) = L + 1 GEMM launches between a begin and an accept kernel:
//   dense_begin_kernel     p ~ N(0, I) (Philox or injected), ke0, delta = q - mu
//   dense_gemm_kick_kernel z = delta . P; epilogue: kick p -= c z, NEXT drift delta' = delta + eps p (other buffer),
//                          and at the two trajectory ends logp = c - 1/2 z.delta, ke = 1/2 |p|^2
//   dense_accept_kernel    Hamiltonian, Metropolis accept, q <- delta + mu, sample row -> [chain, slot, :]
//
// dense_gemm_kick_kernel — one CTA per 128 chains, clusters of 2 CTAs, 15 warps:
//   warp 0      TMA loads of the raw f32 delta tile [128 x 32] (ring of 3; [128 x 16], ring of 4 in TF32 mode)
//   warp 1      TMA loads of P_hi / P_lo [256 x 32 halves]: each CTA of the cluster loads half of every tile and multicasts it
//   warp 2      tcgen05.mma issue (one elected lane; 128 x 256 x 16, accumulators double-buffered over the 512 TMEM cols)
//   warps 3-6   split the raw tile into hi / lo operand tiles in shared memory (element-wise, so the TMA swizzle
//               is preserved) — the split never touches HBM
//   warps 7-14  epilogue: tcgen05.ld, transpose through shared memory, row-contiguous global updates
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kernels.h"
#include "philox.cuh"

namespace gm {

namespace {

constexpr int kTileM = 128;      // chains per CTA
constexpr int kTileN = 256;      // accumulator columns per MMA
#ifndef GM_TC_F16
#define GM_TC_F16 1              // 1: FP16 x 3 split (kind::f16, twice the TF32 MMA rate); 0: TF32 x 3 split (kind::tf32)
#endif
constexpr bool kF16 = (GM_TC_F16 != 0);
// Operand rows are 64 bytes in both modes (SWIZZLE_64B): 32 halves or 16 tf32 words per K stage, two MMAs per stage.
constexpr int kTileK = kF16 ? 32 : 16;   // elements per K stage
constexpr int kUmmaK = kF16 ? 16 : 8;    // elements per tcgen05.mma
constexpr int kOpElem = kF16 ? 2 : 4;    // bytes per operand element
#ifndef GM_TC_STAGES
#define GM_TC_STAGES 3
#endif
#ifndef GM_TC_RAW_STAGES
#define GM_TC_RAW_STAGES (GM_TC_F16 ? 3 : 4)
#endif
constexpr int kStages = GM_TC_STAGES;         // operand stages (A_hi, A_lo, P_hi, P_lo)
constexpr int kRawStages = GM_TC_RAW_STAGES;  // raw f32 delta tiles in flight ahead of the split (16 KB / 8 KB each)
constexpr int kKPadUnit = 32;    // K is padded to a multiple of 32 floats in both layouts
#ifndef GM_TC_CLUSTER
#define GM_TC_CLUSTER 2          // CTAs per cluster sharing every P tile through TMA multicast (1 = no cluster)
#endif
constexpr int kCluster = GM_TC_CLUSTER;
#ifndef GM_TC_EPI_PREFETCH
#define GM_TC_EPI_PREFETCH 0
#endif
#ifndef GM_TC_EPI_PIPE
#define GM_TC_EPI_PIPE 0      // 1: fast epilogue blocks keep the p / delta loads of both 16-row halves in flight across the TMEM read and
                              // refill them block to block (see the epilogue) — measured SLOWER (19.1 vs 16.4 ms per transition): at the 128
                              // registers a 480-thread CTA allows, the 64 load registers spill and the spill stores wait for the loads
#endif
#ifndef GM_TC_EPI_WARPS
#define GM_TC_EPI_WARPS 8
#endif
constexpr int kEpiWarps = GM_TC_EPI_WARPS;   // multiple of 4: kEpiWarps / 4 warps per TMEM lane quarter, interleaved over the 32-column blocks
constexpr int kTrFloats = GM_TC_EPI_PIPE ? 32 * 32 : 16 * 33;   // per epilogue warp: the accumulator rows in transit (a 32 x 32 block, or 16 rows of 33)
constexpr int kCvtWarps = 4;
constexpr int kFirstCvtWarp = 3, kFirstEpiWarp = kFirstCvtWarp + kCvtWarps;
constexpr int kGemmThreads = 32 * (kFirstEpiWarp + kEpiWarps);
constexpr uint32_t kABytes = kTileM * kTileK * kOpElem;   // 8 KB
constexpr uint32_t kBBytes = kTileN * kTileK * kOpElem;   // 16 KB
constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;   // 48 KB
constexpr uint32_t kRawBytes = kTileM * kTileK * 4;       // raw f32 tile: 16 KB (128-byte rows, SWIZZLE_128B) / 8 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifndef GM_TC_WAIT_HINT_NS
#define GM_TC_WAIT_HINT_NS 2000   // suspend-time hint of mbarrier.try_wait: the waiting warp is parked by the hardware
#endif
// Blocking wait on an mbarrier phase.  try_wait with a suspend-time hint parks the warp instead of spinning (the polling
// loop of the producer / consumer warps was 20.9 % of all executed warp-instructions and took issue slots from the
// epilogue warps of the same SM sub-partition); after repeated time-outs the wait backs off with nanosleep.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t spins = 0;
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"((uint32_t)GM_TC_WAIT_HINT_NS)
        : "memory");
    if (done) return;
    if (++spins > 64u) __nanosleep(64);
    if (spins > (1u << 24)) __trap();   // a lost arrival must fault, never hang the GPU
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
#if GM_TC_F16
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
#else
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
#endif
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor: K-major operand tile, rows of kTileK floats, swizzle width = row width (8-row atoms)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);        // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                            // leading byte offset (unused: K extent = one swizzle row)
  constexpr uint32_t kRowBytes = kTileK * kOpElem;   // 64: SWIZZLE_64B
  d |= (uint64_t)(((8 * kRowBytes) >> 4) & 0x3fff) << 32;   // stride byte offset between 8-row groups, bits [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version (sm_100), bits [46,48)
  d |= (uint64_t)4 << 61;                            // layout type SWIZZLE_64B, bits [61,64)
  return d;
}

// instruction descriptor: D f32 (bits 4-5 = 1), A / B format (bits 7-9 / 10-12: 0 = f16, 2 = tf32), both K-major, M = 128, N = 256
constexpr uint32_t kOpFmt = kF16 ? 0u : 2u;
constexpr uint32_t kIdesc = (1u << 4) | (kOpFmt << 7) | (kOpFmt << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

#if !GM_TC_F16
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
#endif

struct GemmArgs {
  int d, kpad, npad;
  size_t n_chains;
  float* p;               // [C, d]
  const float* dl;        // [C, kpad] delta read by this GEMM (also the A operand, via TMA)
  float* dl_next;         // [C, kpad] delta after the next drift, or null
  float coef;             // kick: p -= coef * z   (grad = -z)
  float drift_eps;
  float norm_const;
  float zscale;           // the accumulator holds z / (zscale * dscale[1]): P is pre-scaled by a power of two in FP16 mode
  const float* dscale;    // device [2]: power-of-two scale applied to delta before the FP16 split, and its inverse
  int persist;            // 1: the grid is a set of persistent clusters walking (row-tile group, column chunk) units
  int row_base;           // first chain of this launch's chain block: row coordinate offset of the delta TMA loads (the other
                          // pointers are pre-offset by the host)
  // trajectory ends: per-unit partial row sums, one slot per (column chunk, epilogue warp of the quarter), summed in fixed
  // order by the consumer (dense_accept_kernel) — every launch can then run the persistent unit schedule
  float* quad_part;       // [n_chunks * kEpiWarps / 4][part_ld] or null: zs * sum_cols z * delta
  float* ke_part;         // same shape or null: sum_cols p_new^2
  size_t part_ld;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
dense_gemm_kick_kernel(const __grid_constant__ CUtensorMap map_dl, const __grid_constant__ CUtensorMap map_bhi,
                       const __grid_constant__ CUtensorMap map_blo, const GemmArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* raw_base = base + (size_t)kStages * kStageBytes;
  uint64_t* bars = (uint64_t*)(raw_base + (size_t)kRawStages * kRawBytes);
  uint64_t* b_full = bars;                         // [kStages]    P tiles landed (TMA tx)
  uint64_t* a_full = b_full + kStages;             // [kStages]    hi / lo tiles written by the split warps
  uint64_t* empty = a_full + kStages;              // [kStages]    stage released by the MMAs of BOTH CTAs
  uint64_t* raw_full = empty + kStages;            // [kRawStages] raw delta tile landed
  uint64_t* raw_empty = raw_full + kRawStages;     // [kRawStages] raw tile consumed by the split warps
  uint64_t* tfull = raw_empty + kRawStages;        // [2] accumulator complete
  uint64_t* tempty = tfull + 2;                    // [2] accumulator drained
  uint32_t* tmem_slot = (uint32_t*)(tempty + 2);
  float* tr_base = reinterpret_cast<float*>(raw_base + (size_t)kRawStages * kRawBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = a.npad / kTileN;
  const int k_chunks = a.kpad / kTileK;
  // Work units = (group of kCluster adjacent row tiles, column chunk), chunk fastest.  persist = 0: cluster c owns group c
  // and walks its n_chunks units (needed at the trajectory ends, where the row sums run over all columns).  persist = 1:
  // the grid is one cluster per SM pair; cluster c takes units c, c + W, c + 2W, ... — 512 row tiles on 148 SMs are 3.46
  // waves of whole tiles but 13.8 rounds of units.
  const int cluster_id = (int)blockIdx.x / kCluster, n_clusters = (int)gridDim.x / kCluster;
  const int my_rank = (int)blockIdx.x - cluster_id * kCluster;
  const int n_groups = (int)((a.n_chains + (size_t)kTileM * kCluster - 1) / ((size_t)kTileM * kCluster));
  const int n_units_all = n_groups * n_chunks;
  const int n_units = a.persist ? (cluster_id < n_units_all ? (n_units_all - cluster_id + n_clusters - 1) / n_clusters : 0) : n_chunks;
  auto unit_of = [&](int j, int& m0, int& n) {
    const int u = a.persist ? cluster_id + j * n_clusters : cluster_id * n_chunks + j;
    const int grp = u / n_chunks;
    n = u - grp * n_chunks;
    m0 = (grp * kCluster + my_rank) * kTileM;
  };
  const int n_iters = n_units * k_chunks;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&a_full[i], kCvtWarps); mbar_init(&empty[i], kCluster); }
    for (int i = 0; i < kRawStages; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], kCvtWarps); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kCluster > 1) cluster_sync_all();     // peer barriers are initialised before any remote arrival
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = kCluster > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = (uint16_t)((1u << kCluster) - 1u);

  if (warp == 0) {
    if (lane == 0) {
      // ===== raw delta tiles (re-streamed once per column chunk; the 512 KB row block stays L2-resident)
      for (int it = 0; it < n_iters; ++it) {
        const int r = it % kRawStages;
        mbar_wait(&raw_empty[r], ((uint32_t)(it / kRawStages) & 1u) ^ 1u);
        mbar_expect_tx(&raw_full[r], kRawBytes);
        int m0, n;
        unit_of(it / k_chunks, m0, n);
        tma_load_2d(raw_base + (size_t)r * kRawBytes, &map_dl, &raw_full[r], (it % k_chunks) * kTileK, m0 + a.row_base);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== P tiles: my half of every tile, multicast to the cluster
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % kStages;
        const int k = it % k_chunks;
        int m0, n;
        unit_of(it / k_chunks, m0, n);
        mbar_wait(&empty[s], ((uint32_t)(it / kStages) & 1u) ^ 1u);
        unsigned char* st = base + (size_t)s * kStageBytes + 2 * kABytes;
        mbar_expect_tx(&b_full[s], 2 * kBBytes);
        if constexpr (kCluster > 1) {
          const int half_rows = kTileN / kCluster;
          const uint32_t off = crank * (uint32_t)(half_rows * kTileK * kOpElem);
          tma_load_2d_mc(st + off, &map_bhi, &b_full[s], k * kTileK, n * kTileN + (int)crank * half_rows, kMask);
          tma_load_2d_mc(st + kBBytes + off, &map_blo, &b_full[s], k * kTileK, n * kTileN + (int)crank * half_rows, kMask);
        } else {
          tma_load_2d(st, &map_bhi, &b_full[s], k * kTileK, n * kTileN);
          tma_load_2d(st + kBBytes, &map_blo, &b_full[s], k * kTileK, n * kTileN);
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ===== MMA issuer
      int it = 0;
      for (int j = 0; j < n_units; ++j) {
        const int acc = j & 1;
        mbar_wait(&tempty[acc], (uint32_t)((j >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kTileN);
        for (int k = 0; k < k_chunks; ++k, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (uint32_t)(it / kStages) & 1u;
          mbar_wait(&a_full[s], ph);
          mbar_wait(&b_full[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(base + (size_t)s * kStageBytes);
          const uint64_t d_ahi = umma_desc(st), d_alo = umma_desc(st + kABytes);
          const uint64_t d_bhi = umma_desc(st + 2 * kABytes), d_blo = umma_desc(st + 2 * kABytes + kBBytes);
#pragma unroll
          for (int kk = 0; kk < kTileK / kUmmaK; ++kk) {
            const uint64_t adv = (uint64_t)((kk * kUmmaK * kOpElem) >> 4);   // 32 bytes per K step inside the swizzle row
            tc_mma_tf32(tmem_d, d_alo + adv, d_bhi + adv, kIdesc, (k | kk) ? 1u : 0u);   // small terms first
            tc_mma_tf32(tmem_d, d_ahi + adv, d_blo + adv, kIdesc, 1u);
            tc_mma_tf32(tmem_d, d_ahi + adv, d_bhi + adv, kIdesc, 1u);
          }
          if constexpr (kCluster > 1) tc_commit_mc(&empty[s], kMask);   // both CTAs' producers wait for both MMAs
          else tc_commit(&empty[s]);
        }
        tc_commit(&tfull[acc]);           // accumulator complete
      }
    }
  } else if (warp < kFirstEpiWarp) {
    // ===== operand split.  TF32 mode: raw f32 tile -> hi = rna_tf32(x), lo = rna_tf32(x - hi) at the same (swizzled) byte
    // offsets.  FP16 mode: hi = half(x), lo = half(x - hi); the raw tile has 128-byte rows under SWIZZLE_128B (16-byte
    // chunk c of row r sits at chunk c ^ (r & 7)), the operand tiles 64-byte rows under SWIZZLE_64B (chunk c of row r at
    // c ^ ((r >> 1) & 3)); a thread turns 8 consecutive floats of a row into one 16-byte chunk of each operand tile.
    const int t = (warp - kFirstCvtWarp) * 32 + lane;          // 0..127
    const float dl_scale = kF16 ? a.dscale[0] : 1.f;
    (void)dl_scale;
    for (int it = 0; it < n_iters; ++it) {
      const int r = it % kRawStages, s = it % kStages;
      mbar_wait(&raw_full[r], (uint32_t)(it / kRawStages) & 1u);
      mbar_wait(&empty[s], ((uint32_t)(it / kStages) & 1u) ^ 1u);
#if GM_TC_F16
      const unsigned char* src = raw_base + (size_t)r * kRawBytes;
      unsigned char* dhi = base + (size_t)s * kStageBytes;
      unsigned char* dlo = dhi + kABytes;
      const float sa = dl_scale;
#pragma unroll
      for (int i = 0; i < (kTileM * 4) / (32 * kCvtWarps); ++i) {
        const int u = t + i * 32 * kCvtWarps;
        const int row = u >> 2, oc = u & 3, sw = row & 7;
        const float4 x0 = *reinterpret_cast<const float4*>(src + row * 128 + (((2 * oc) ^ sw) << 4));
        const float4 x1 = *reinterpret_cast<const float4*>(src + row * 128 + (((2 * oc + 1) ^ sw) << 4));
        const float xs[8] = {x0.x * sa, x0.y * sa, x0.z * sa, x0.w * sa, x1.x * sa, x1.y * sa, x1.z * sa, x1.w * sa};
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // packed conversions (F2FP.F16.F32.PACK_AB: one FMA-pipe instruction per pair, not the quarter-rate XU F2F)
          const __half2 h = __floats2half2_rn(xs[2 * e], xs[2 * e + 1]);
          const float2 hf = __half22float2(h);
          const __half2 l = __floats2half2_rn(xs[2 * e] - hf.x, xs[2 * e + 1] - hf.y);
          hw[e] = *reinterpret_cast<const uint32_t*>(&h);
          lw[e] = *reinterpret_cast<const uint32_t*>(&l);
        }
        const int so = row * 64 + ((oc ^ ((row >> 1) & 3)) << 4);
        *reinterpret_cast<uint4*>(dhi + so) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(dlo + so) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
#else
      const float4* src = reinterpret_cast<const float4*>(raw_base + (size_t)r * kRawBytes);
      float4* dhi = reinterpret_cast<float4*>(base + (size_t)s * kStageBytes);
      float4* dlo = reinterpret_cast<float4*>(base + (size_t)s * kStageBytes + kABytes);
#pragma unroll
      for (int i = 0; i < (int)(kRawBytes / 16) / (32 * kCvtWarps); ++i) {
        const float4 x = src[t + i * 32 * kCvtWarps];
        float4 h, l;
        h.x = tf32_rna(x.x); h.y = tf32_rna(x.y); h.z = tf32_rna(x.z); h.w = tf32_rna(x.w);
        l.x = tf32_rna(x.x - h.x); l.y = tf32_rna(x.y - h.y); l.z = tf32_rna(x.z - h.z); l.w = tf32_rna(x.w - h.w);
        dhi[t + i * 32 * kCvtWarps] = h;
        dlo[t + i * 32 * kCvtWarps] = l;
      }
#endif
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) { mbar_arrive(&a_full[s]); mbar_arrive(&raw_empty[r]); }
    }
  } else {
    // ===== epilogue.  tcgen05.ld hands every thread one accumulator ROW (32 columns at a time); 16 rows at a time
    // are transposed through shared memory so that global accesses run along rows: lane <-> column, one 128-byte
    // line per warp instruction for p and delta.
    const int ew = warp - kFirstEpiWarp;
    const int q4 = warp & 3;                     // TMEM lane quarter this warp may access
    const float zs = kF16 ? a.zscale * a.dscale[1] : a.zscale;   // z = zs * accumulator
    const float coef = a.coef * zs;
    float* tr = tr_base + (size_t)ew * kTrFloats;
    // the kEpiWarps / 4 warps that share a quarter interleave over the 32-column blocks
    int cb_first = 0;
    for (int w = kFirstEpiWarp; w < warp; ++w) cb_first += ((w & 3) == q4) ? 1 : 0;
    size_t row0 = 0;
    int nrows = 0;
    float quad = 0.f, ke = 0.f;                  // lane r holds the sums of row r (persist = 0: one row tile per CTA)
    for (int j = 0; j < n_units; ++j) {
      int m0, n;
      unit_of(j, m0, n);
      row0 = (size_t)m0 + (size_t)q4 * 32;
      nrows = row0 < a.n_chains ? (int)((a.n_chains - row0) < 32 ? (a.n_chains - row0) : 32) : 0;
      const int acc = j & 1;
      // Fast blocks (every launch but the trajectory ends, every full 32 x 32 block of a d % 4 == 0 target): no bounds tests,
      // no row sums, 128-bit global accesses — a lane owns 4 consecutive columns of 4 rows per 16-row half, one warp instruction
      // moves four whole 128-byte lines.  (GM_TC_EPI_PREFETCH = 1 keeps the p / delta loads of the NEXT half in flight while the
      // current half is transposed, updated and stored — measured SLOWER, 17.9 vs 16.4 ms per transition: the registers it
      // takes spill, and the epilogue is not short of loads in flight: tensor pipe 57 %, L1 52 %, crossbar 40 %, DRAM 33 %, at
      // a power-limited 1.54 GHz, nothing saturated.)
      const bool unit_fast = a.dl_next && !a.quad_part && !a.ke_part && nrows == 32 && (a.d & 3) == 0;
      auto blk_fast = [&](int cb) { return unit_fast && cb < kTileN / 32 && n * kTileN + cb * 32 + 32 <= a.d; };
      const int sub = lane >> 3, ch = lane & 7;
      const uint32_t d4 = 4u * (uint32_t)a.d, k4 = 4u * (uint32_t)a.kpad;
      auto issue = [&](int cb, int rb, float4 (&pv4)[4], float4 (&dv4)[4]) {
        const size_t r_first = row0 + rb + sub;
        const int cc = n * kTileN + cb * 32 + 4 * ch;
        const float* p_ptr = a.p + r_first * (size_t)a.d + cc;
        const float* dl_ptr = a.dl + r_first * (size_t)a.kpad + cc;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
#ifdef GM_TC_EXPERIMENT_NOLOAD
          pv4[jj] = make_float4(1.f, 1.f, 1.f, 1.f); dv4[jj] = make_float4(2.f, 2.f, 2.f, 2.f);
#else
          pv4[jj] = __ldcs(reinterpret_cast<const float4*>(p_ptr + jj * d4));
          dv4[jj] = *reinterpret_cast<const float4*>(dl_ptr + jj * k4);
#endif
        }
      };
      float4 pA[4], dA[4], pB[4], dB[4];
#if GM_TC_EPI_PREFETCH
      bool preA = blk_fast(cb_first);
      if (preA) issue(cb_first, 0, pA, dA);
#else
      bool preA = false;
#endif
#if GM_TC_EPI_PIPE
      // The epilogue is what the tensor pipe waits for (both accumulator buffers full), and its time was the global-load
      // latency of p / delta, exposed twice per 32 x 32 block (long_scoreboard 63 % of the kernel's stall samples).  Pipelined
      // form: the loads of BOTH 16-row halves of a block are issued before the accumulator is complete / read from TMEM, the
      // block goes to shared memory in two 16-column reads (so the accumulator registers never coexist with more than the 64
      // load registers), and each half's registers are refilled with the next block's loads as soon as the half is stored.
      // this warp's blocks are cb_first, cb_first + step, ...; the leading n_fast of them take the pipelined path
      constexpr int kStep = kEpiWarps / 4;
      int n_fast = 0;
      for (int cb = cb_first; cb < kTileN / 32 && blk_fast(cb); cb += kStep) ++n_fast;
      if (n_fast > 0) { issue(cb_first, 0, pA, dA); issue(cb_first, 16, pB, dB); }
      auto process = [&](int cb, int rb, const float4 (&pv4)[4], const float4 (&dv4)[4]) {
        const int c0 = n * kTileN + cb * 32;
        const size_t r_first = row0 + rb + sub;
        float* p_ptr = a.p + r_first * (size_t)a.d + c0 + 4 * ch;
        float* dn_ptr = a.dl_next + r_first * (size_t)a.kpad + c0 + 4 * ch;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int r = rb + 4 * jj + sub;
          const float4 zv = *reinterpret_cast<const float4*>(tr + r * 32 + ((ch ^ (r & 7)) << 2));
          float4 pn, dn;
          pn.x = fmaf(-coef, zv.x, pv4[jj].x); pn.y = fmaf(-coef, zv.y, pv4[jj].y);
          pn.z = fmaf(-coef, zv.z, pv4[jj].z); pn.w = fmaf(-coef, zv.w, pv4[jj].w);
          dn.x = fmaf(a.drift_eps, pn.x, dv4[jj].x); dn.y = fmaf(a.drift_eps, pn.y, dv4[jj].y);
          dn.z = fmaf(a.drift_eps, pn.z, dv4[jj].z); dn.w = fmaf(a.drift_eps, pn.w, dv4[jj].w);
          __stcs(reinterpret_cast<float4*>(p_ptr + jj * d4), pn);
          *reinterpret_cast<float4*>(dn_ptr + jj * k4) = dn;
        }
      };
#endif
      mbar_wait(&tfull[acc], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
#pragma unroll 1
      for (int cb = cb_first; cb < kTileN / 32; cb += kEpiWarps / 4) {
#if GM_TC_EPI_PIPE
        if ((cb - cb_first) / kStep < n_fast) {
          const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * kTileN + cb * 32);
          // row `lane` of the block as eight 16-byte chunks, chunk i at position i ^ (lane & 7): conflict-free for these
          // row-wise writes and for the 4-rows-by-8-chunks reads of process()
          float4* trow = reinterpret_cast<float4*>(tr + lane * 32);
          __syncwarp();                          // the previous block's reads of the staging buffer are done
          {
            float z[16];
            tc_ld16(taddr, z);
#pragma unroll
            for (int i = 0; i < 4; ++i) trow[i ^ (lane & 7)] = make_float4(z[4 * i], z[4 * i + 1], z[4 * i + 2], z[4 * i + 3]);
          }
          {
            float z[16];
            tc_ld16(taddr + 16u, z);
#pragma unroll
            for (int i = 0; i < 4; ++i) trow[(4 + i) ^ (lane & 7)] = make_float4(z[4 * i], z[4 * i + 1], z[4 * i + 2], z[4 * i + 3]);
          }
          __syncwarp();
          process(cb, 0, pA, dA);
          // the next fast block's loads refill each half's registers as soon as the half is stored; after the last one the
          // same block is loaded once more (unused) so that the loads stay unconditional and the 64 values stay in registers
          const int cbn = ((cb - cb_first) / kStep + 1 < n_fast) ? cb + kStep : cb;
          issue(cbn, 0, pA, dA);
          process(cb, 16, pB, dB);
          issue(cbn, 16, pB, dB);
          continue;
        }
#endif
        float z[32];
        tc_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * kTileN + cb * 32), z);
        const int c0 = n * kTileN + cb * 32;
        if (c0 >= a.d) continue;                 // padded columns (warp-uniform)
#ifdef GM_TC_EXPERIMENT_NOEPI   // timing experiments (wrong results): -DGM_TC_EXPERIMENT_NOEPI / _NOLOAD / _NOSTORE, see DESIGN.md K3
        if (z[0] != 123456.f) continue;
#endif
        const int col = c0 + lane;
        const bool col_ok = col < a.d;
#if !GM_TC_EPI_PIPE
        if (blk_fast(cb)) {
          // one 16-row half: transpose staging (row L of the half as eight 16-byte chunks, chunk i at position i ^ (L & 7):
          // conflict-free for the row-wise writes and for the 4-rows-by-8-chunks reads), kick, next drift, stores
          auto half = [&](int rb, const float4 (&pv4)[4], const float4 (&dv4)[4]) {
            __syncwarp();
            if ((lane & 16) == rb) {
              const int L = lane & 15;
              float4* trow = reinterpret_cast<float4*>(tr + L * 32);
#pragma unroll
              for (int i = 0; i < 8; ++i) trow[i ^ (L & 7)] = make_float4(z[4 * i], z[4 * i + 1], z[4 * i + 2], z[4 * i + 3]);
            }
            __syncwarp();
            const size_t r_first = row0 + rb + sub;
            float* p_ptr = a.p + r_first * (size_t)a.d + c0 + 4 * ch;
            float* dn_ptr = a.dl_next + r_first * (size_t)a.kpad + c0 + 4 * ch;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int r = 4 * jj + sub;
              const float4 zv = *reinterpret_cast<const float4*>(tr + r * 32 + ((ch ^ (r & 7)) << 2));
              float4 pn, dn;
              pn.x = fmaf(-coef, zv.x, pv4[jj].x); pn.y = fmaf(-coef, zv.y, pv4[jj].y);
              pn.z = fmaf(-coef, zv.z, pv4[jj].z); pn.w = fmaf(-coef, zv.w, pv4[jj].w);
              dn.x = fmaf(a.drift_eps, pn.x, dv4[jj].x); dn.y = fmaf(a.drift_eps, pn.y, dv4[jj].y);
              dn.z = fmaf(a.drift_eps, pn.z, dv4[jj].z); dn.w = fmaf(a.drift_eps, pn.w, dv4[jj].w);
#ifdef GM_TC_EXPERIMENT_NOSTORE
              if (pn.x == 123456.f)
#endif
              {
                __stcs(reinterpret_cast<float4*>(p_ptr + jj * d4), pn);
                *reinterpret_cast<float4*>(dn_ptr + jj * k4) = dn;
              }
            }
          };
#if GM_TC_EPI_PREFETCH
          if (!preA) issue(cb, 0, pA, dA);
          issue(cb, 16, pB, dB);                         // second half in flight under the first
          half(0, pA, dA);
          const int cbn = cb + kEpiWarps / 4;
          preA = blk_fast(cbn);
          if (preA) issue(cbn, 0, pA, dA);               // next block's first half in flight under the second
          half(16, pB, dB);
#else
          issue(cb, 0, pA, dA);
          half(0, pA, dA);
          issue(cb, 16, pB, dB);
          half(16, pB, dB);
#endif
          continue;
        }
#endif
        preA = false;
#pragma unroll 1
        for (int rb = 0; rb < 32; rb += 16) {
          __syncwarp();
          if ((lane & 16) == rb) {
#pragma unroll
            for (int c = 0; c < 32; ++c) tr[(lane & 15) * 33 + c] = z[c];
          }
          __syncwarp();
          // block pointers once per 16 rows; the rows then advance by d / kpad elements
          float* p_blk = a.p + (row0 + rb) * (size_t)a.d + col;
          const float* dl_blk = a.dl + (row0 + rb) * (size_t)a.kpad + col;
          const uint32_t d32 = (uint32_t)a.d, k32 = (uint32_t)a.kpad;
          float pv[16], dv[16];
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) {
            const int r = rb + rr;
            pv[rr] = 0.f; dv[rr] = 0.f;
            if (r < nrows && col_ok) {
              pv[rr] = __ldcs(p_blk + rr * d32);
              dv[rr] = dl_blk[rr * k32];
            }
          }
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) {
            const int r = rb + rr;
            if (r < nrows) {
              const float zv = tr[rr * 33 + lane];
              float pn = 0.f;
              if (col_ok) {
                pn = fmaf(-coef, zv, pv[rr]);
                __stcs(p_blk + rr * d32, pn);
                if (a.dl_next) a.dl_next[(row0 + r) * (size_t)a.kpad + col] = fmaf(a.drift_eps, pn, dv[rr]);
              }
              if (a.quad_part || a.ke_part) {        // trajectory ends only: row sums over the 32 columns
                float s1 = zv * dv[rr], s2 = pn * pn;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
                if (lane == r) { quad += s1; ke += s2; }
              }
            }
          }
        }
      }
      if (a.quad_part || a.ke_part) {
        // this warp's partial row sums of the unit (its column blocks of chunk n): one slot per (chunk, warp of the quarter)
        const size_t slot = (size_t)(n * (kEpiWarps / 4) + cb_first);
        if (lane < nrows) {
          if (a.quad_part) a.quad_part[slot * a.part_ld + row0 + lane] = zs * quad;
          if (a.ke_part) a.ke_part[slot * a.part_ld + row0 + lane] = ke;
        }
        quad = 0.f; ke = 0.f;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kCluster > 1) cluster_sync_all();     // no CTA exits while its peer can still multicast into it
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- elementwise kernels -----------------------------------------------------------------------------
struct BeginArgs {
  size_t n_chains; int d, kpad;
  unsigned long long chain_offset; PhiloxKey key; uint32_t step;
  const float* q; float* p; const float* mu; float* dl; float* ke0;
  const float* inj_normals;   // [C, d] for this transition or null
  unsigned int* dmax_bits;    // [1] max |delta| of the batch as float bits (non-negative floats order like unsigned ints)
};

// one warp per chain: momentum, ke0, delta = q - mu (padded columns zero)
__global__ void __launch_bounds__(256) dense_begin_kernel(const BeginArgs a) {
  const size_t chain = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (chain >= a.n_chains) return;
  const unsigned long long gchain = a.chain_offset + chain;
  float ke = 0.f, dmax = 0.f;
  for (int b = lane; b * 4 < a.kpad; b += 32) {
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (b * 4 < a.d) {
      if (a.inj_normals) {
        for (int k = 0; k < 4; ++k)
          if (b * 4 + k < a.d) z[k] = a.inj_normals[chain * a.d + b * 4 + k];
      } else {
        normals_from_block<false>(philox4x32_10(philox_ctr(gchain, a.step, 0u, (uint32_t)b), a.key), z);
      }
    }
    float dl[4];
    for (int k = 0; k < 4; ++k) {
      const int c = b * 4 + k;
      dl[k] = 0.f;
      if (c < a.d) {
        a.p[chain * a.d + c] = z[k];
        ke += z[k] * z[k];
        dl[k] = a.q[chain * a.d + c] - a.mu[c];
        const float ad = fabsf(dl[k]);
        if (ad < INFINITY) dmax = fmaxf(dmax, ad);
      }
    }
    *reinterpret_cast<float4*>(a.dl + chain * a.kpad + b * 4) = make_float4(dl[0], dl[1], dl[2], dl[3]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    ke += __shfl_xor_sync(0xffffffffu, ke, o);
    dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  }
  if (lane == 0) {
    a.ke0[chain] = 0.5f * ke;
    if (a.dmax_bits && dmax > 0.f) atomicMax(a.dmax_bits, __float_as_uint(dmax));
  }
}

// max |delta| of the batch -> the power-of-two operand scale of this transition's GEMMs (2^8 / max, rounded down to a power
// of two: 2^7 headroom below FP16's largest finite value for the trajectory); resets the maximum for the next transition
__global__ void dense_scale_kernel(unsigned int* dmax_bits, float* dscale) {
  const float m = __uint_as_float(*dmax_bits);
  int e = 0;
  if (m > 0.f && m < INFINITY) {
    int ex;
    frexpf(m, &ex);              // m = f * 2^ex, f in [0.5, 1)
    e = 8 - ex;
  }
  e = e < -100 ? -100 : (e > 100 ? 100 : e);
  dscale[0] = ldexpf(1.f, e);
  dscale[1] = ldexpf(1.f, -e);
  *dmax_bits = 0u;
}

struct AcceptArgs {
  size_t n_chains; int d, kpad;
  unsigned long long chain_offset; PhiloxKey key; uint32_t step;
  float* q; const float* dl; const float* mu; const float* p;
  // log densities / kinetic energy of the two trajectory ends as partial row sums [n_part][part_ld] of the GEMM epilogues
  const float* quad0; const float* quad1; const float* ke0; const float* ke1_part;
  int n_part; size_t part_ld; float norm_const;
  float* out; size_t out_n; long long slot;        // slot < 0: not recorded
  unsigned long long* accept_total; unsigned long long* diverge_total;
  const float* inj_lnu;      // [C] for this transition or null
  float* diag_logacc; uint8_t* diag_acc; float* diag_pq; float* diag_pp;   // per-transition slices or null
};

// one warp per chain: Metropolis accept (batched_hmc.rs:148-162), q <- delta + mu, sample write-out
__global__ void __launch_bounds__(256) dense_accept_kernel(const AcceptArgs a) {
  const size_t chain = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (chain >= a.n_chains) return;
  float q0 = 0.f, q1 = 0.f, k1 = 0.f;           // fixed summation order over the (chunk, warp) slots: deterministic
  for (int i = 0; i < a.n_part; ++i) {
    q0 += a.quad0[(size_t)i * a.part_ld + chain];
    q1 += a.quad1[(size_t)i * a.part_ld + chain];
    k1 += a.ke1_part[(size_t)i * a.part_ld + chain];
  }
  const float logp0 = a.norm_const - 0.5f * q0, logp1 = a.norm_const - 0.5f * q1, ke1 = 0.5f * k1;
  const float log_accept = (logp1 - logp0) + (a.ke0[chain] - ke1);
  float ln_u;
  if (a.inj_lnu) ln_u = a.inj_lnu[chain];
  else ln_u = logf(u01(philox4x32_10(philox_ctr(a.chain_offset + chain, a.step, 1u, 0u), a.key).x));
  const bool accept = ln_u <= log_accept;
  const bool finite = (log_accept == log_accept) && (fabsf(log_accept) < INFINITY);
  float* qrow = a.q + chain * a.d;
  const float* drow = a.dl + chain * a.kpad;
  for (int c = lane; c < a.d; c += 32) {
    const float prop = drow[c] + a.mu[c];
    const float v = accept ? prop : qrow[c];
    if (accept) qrow[c] = v;
    if (a.slot >= 0) __stcs(a.out + (chain * a.out_n + (size_t)a.slot) * a.d + c, v);
    if (a.diag_pq) { a.diag_pq[chain * a.d + c] = prop; a.diag_pp[chain * a.d + c] = a.p[chain * a.d + c]; }
  }
  if (lane == 0) {
    if (accept) atomicAdd(a.accept_total, 1ull);
    if (!finite) atomicAdd(a.diverge_total, 1ull);
    if (a.diag_logacc) { a.diag_logacc[chain] = log_accept; a.diag_acc[chain] = accept ? 1 : 0; }
  }
}

// ---- host side -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D tensor [rows, cols] (cols contiguous) of f32 (elem = 4) or f16 (elem = 2), box [box_rows, kTileK elements]; the
// swizzle width equals the box row: 64 bytes (operands; raw tile in TF32 mode) or 128 bytes (raw f32 tile in FP16 mode)
bool make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows, int elem) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * (cuuint64_t)elem};
  cuuint32_t box[2] = {(cuuint32_t)kTileK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const bool wide = (kTileK * elem == 128);
  return enc(m, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides,
             box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline float host_tf32_rna(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u += 0x1000u;            // round to nearest, ties away from zero (cvt.rna)
  u &= 0xffffe000u;
  float r;
  std::memcpy(&r, &u, 4);
  return r;
}

}  // namespace

struct DenseTc {
  int d = 0, kpad = 0, npad = 0;
  size_t n_chains = 0;
  float *p = nullptr, *dl[2] = {nullptr, nullptr}, *mu = nullptr;
  void *b_hi = nullptr, *b_lo = nullptr;   // P split into hi / lo operand arrays [npad, kpad]: tf32-in-f32 or f16
  float zscale = 1.f;                      // 1 / (power-of-two scale applied to P before the FP16 split)
  float* dscale = nullptr;                 // device [2]: power-of-two scale of delta for this transition, and its inverse
  unsigned int* dmax_bits = nullptr;       // device [1]: max |delta| of the batch (float bits)
  int sms = 0;                             // SMs of the device (persistent grid size)
  float *ke0 = nullptr;
  float* part = nullptr;                   // [3][n_part][C]: partial row sums of the two trajectory-end GEMMs (quad0, quad1, ke1)
  int n_part = 0;
  size_t block_chains = 0;                 // chains per L2-resident block (0: all chains in one block)
  float norm_const = 0.f;
  CUtensorMap map_dl[2], map_bhi, map_blo;
  size_t smem = 0;
};

void dense_tc_destroy(DenseTc* t) {
  if (!t) return;
  cudaFree(t->p); cudaFree(t->dl[0]); cudaFree(t->dl[1]); cudaFree(t->b_hi); cudaFree(t->b_lo); cudaFree(t->mu);
  cudaFree(t->ke0); cudaFree(t->part);
  cudaFree(t->dscale); cudaFree(t->dmax_bits);
  delete t;
}

// params: [mu[d], P[d*d] row-major (symmetric), norm_const] in double
DenseTc* dense_tc_create(size_t n_chains, int d, const double* params, const char** err) {
  static const char* e_alloc = "dense tensor-core path: device allocation failed";
  static const char* e_map = "dense tensor-core path: cuTensorMapEncodeTiled unavailable or failed";
  DenseTc* t = new DenseTc();
  t->d = d; t->n_chains = n_chains;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&t->sms, cudaDevAttrMultiProcessorCount, dev); }
  t->kpad = ((d + kKPadUnit - 1) / kKPadUnit) * kKPadUnit;
  t->npad = ((d + kTileN - 1) / kTileN) * kTileN;
  t->norm_const = (float)params[(size_t)d + (size_t)d * d];
  t->n_part = (t->npad / kTileN) * (kEpiWarps / 4);
  {
    // Chain blocks (experiment, off by default): GMCMC_DENSE_WAVES = w > 0 walks the chains in blocks of w unit waves whose
    // p / delta / delta' working set stays in the 126 MB L2 from one leapfrog to the next, so the epilogue's loads and stores
    // hit L2 instead of HBM.  Measured SLOWER (65,536 chains, d = 1000: 21.1 / 20.2 / 18.9 / 18.2 ms per transition for w = 1 /
    // 2 / 3 / 4 against 16.5 ms unblocked): the epilogue is not bound by HBM latency, and every extra launch pays its own
    // pipeline fill and tail.
    int waves = 0;
    if (const char* w = std::getenv("GMCMC_DENSE_WAVES")) waves = std::atoi(w);
    const int n_chunks = t->npad / kTileN, clusters = t->sms / kCluster;
    if (waves > 0 && clusters > 0) {
      const size_t groups = std::max<size_t>(1, (size_t)waves * clusters / n_chunks);
      t->block_chains = groups * kTileM * kCluster;
    }
  }
  const size_t C = n_chains;
  bool ok = cudaMalloc(&t->p, C * d * 4) == cudaSuccess &&
            cudaMalloc(&t->dl[0], C * (size_t)t->kpad * 4) == cudaSuccess && cudaMalloc(&t->dl[1], C * (size_t)t->kpad * 4) == cudaSuccess &&
            cudaMemset(t->dl[1], 0, C * (size_t)t->kpad * 4) == cudaSuccess &&
            cudaMalloc(&t->b_hi, (size_t)t->npad * t->kpad * kOpElem) == cudaSuccess && cudaMalloc(&t->b_lo, (size_t)t->npad * t->kpad * kOpElem) == cudaSuccess &&
            cudaMalloc(&t->mu, (size_t)d * 4) == cudaSuccess && cudaMalloc(&t->ke0, C * 4) == cudaSuccess &&
            cudaMalloc(&t->part, 3 * (size_t)(t->npad / kTileN) * (kEpiWarps / 4) * C * 4) == cudaSuccess &&
            cudaMalloc(&t->dscale, 8) == cudaSuccess && cudaMalloc(&t->dmax_bits, 4) == cudaSuccess &&
            cudaMemset(t->dmax_bits, 0, 4) == cudaSuccess;
  if (!ok) { *err = e_alloc; dense_tc_destroy(t); return nullptr; }
  // B operand: rows = output column n, cols = k (K-major); P symmetric so B[n][k] = P[k][n] = P[n][k]
  std::vector<float> mu(d);
  for (int i = 0; i < d; ++i) mu[i] = (float)params[i];
  const size_t nel = (size_t)t->npad * t->kpad;
  if (kF16) {
    // FP16 has 5 exponent bits: P is scaled by a power of two so that its largest entry sits near 2^12 (hi never
    // overflows, lo = P s - hi stays normal down to entries 2^-14 times smaller); the epilogue undoes the scale.
    double amax = 0.0;
    for (size_t i = 0; i < (size_t)d * d; ++i) amax = std::max(amax, std::fabs(params[(size_t)d + i]));
    int e = 0;
    if (amax > 0.0) e = 12 - (int)std::ceil(std::log2(amax));
    e = std::max(-100, std::min(100, e));
    const float sc = std::ldexp(1.0f, e);
    t->zscale = std::ldexp(1.0f, -e);
    std::vector<__half> bh(nel, __float2half_rn(0.f)), bl(nel, __float2half_rn(0.f));
    for (int n = 0; n < d; ++n)
      for (int k = 0; k < d; ++k) {
        const float v = (float)params[(size_t)d + (size_t)k * d + n] * sc;
        const __half hi = __float2half_rn(v);
        bh[(size_t)n * t->kpad + k] = hi;
        bl[(size_t)n * t->kpad + k] = __float2half_rn(v - __half2float(hi));
      }
    ok = cudaMemcpy(t->b_hi, bh.data(), nel * 2, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(t->b_lo, bl.data(), nel * 2, cudaMemcpyHostToDevice) == cudaSuccess;
  } else {
    std::vector<float> bh(nel, 0.f), bl(nel, 0.f);
    for (int n = 0; n < d; ++n)
      for (int k = 0; k < d; ++k) {
        const float v = (float)params[(size_t)d + (size_t)k * d + n];
        const float hi = host_tf32_rna(v);
        bh[(size_t)n * t->kpad + k] = hi;
        bl[(size_t)n * t->kpad + k] = host_tf32_rna(v - hi);
      }
    ok = cudaMemcpy(t->b_hi, bh.data(), nel * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(t->b_lo, bl.data(), nel * 4, cudaMemcpyHostToDevice) == cudaSuccess;
  }
  ok = ok && cudaMemcpy(t->mu, mu.data(), (size_t)d * 4, cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok) { *err = e_alloc; dense_tc_destroy(t); return nullptr; }
  ok = make_map(&t->map_dl[0], t->dl[0], C, (uint64_t)t->kpad, kTileM, 4) && make_map(&t->map_dl[1], t->dl[1], C, (uint64_t)t->kpad, kTileM, 4) &&
       make_map(&t->map_bhi, t->b_hi, (uint64_t)t->npad, (uint64_t)t->kpad, kTileN / kCluster, kOpElem) &&
       make_map(&t->map_blo, t->b_lo, (uint64_t)t->npad, (uint64_t)t->kpad, kTileN / kCluster, kOpElem);
  if (!ok) { *err = e_map; dense_tc_destroy(t); return nullptr; }
  t->smem = (size_t)kStages * kStageBytes + (size_t)kRawStages * kRawBytes + 1024 /*alignment slack*/ + 256 /*barriers*/ +
            (size_t)kEpiWarps * kTrFloats * 4 /*epilogue transpose*/;
  if (cudaFuncSetAttribute(dense_gemm_kick_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t->smem) != cudaSuccess) {
    *err = "dense tensor-core path: shared-memory opt-in failed";
    dense_tc_destroy(t);
    return nullptr;
  }
  return t;
}

// GEMM on delta buffer `buf` (0 / 1) for the chain block [c0, c0 + nb); drift_eps != 0 fuses the next leapfrog's drift and
// writes the other buffer; quad_part / ke_part (set of partial slots of this block's chains) at the trajectory ends
static cudaError_t gemm_kick(DenseTc* t, size_t c0, size_t nb, int buf, float coef, float drift_eps, float* quad_part, float* ke_part,
                             cudaStream_t st) {
  GemmArgs g;
  g.d = t->d; g.kpad = t->kpad; g.npad = t->npad; g.n_chains = nb;
  g.p = t->p + c0 * (size_t)t->d; g.dl = t->dl[buf] + c0 * (size_t)t->kpad;
  g.dl_next = drift_eps != 0.f ? t->dl[buf ^ 1] + c0 * (size_t)t->kpad : nullptr;
  g.coef = coef; g.drift_eps = drift_eps; g.norm_const = t->norm_const; g.zscale = t->zscale; g.dscale = t->dscale;
  g.quad_part = quad_part ? quad_part + c0 : nullptr; g.ke_part = ke_part ? ke_part + c0 : nullptr; g.part_ld = t->n_chains;
  g.row_base = (int)c0;
  // persistent clusters over (row-tile group, column chunk) units, chunk fastest, for every launch
  const size_t n_groups = (nb + (size_t)kTileM * kCluster - 1) / ((size_t)kTileM * kCluster);
  const size_t n_units = n_groups * (size_t)(t->npad / kTileN);
  const size_t clusters = std::min<size_t>(n_units, (size_t)std::max(1, t->sms / kCluster));
  g.persist = 1;
  const unsigned blocks = (unsigned)(clusters * kCluster);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = t->smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, dense_gemm_kick_kernel, t->map_dl[buf], t->map_bhi, t->map_blo, g);
}

// One HMC transition.  q: [C, d] current positions (in/out).  Returns the number of kernel launches (or -1).
int dense_tc_transition(DenseTc* t, const DenseTcStep& S, cudaStream_t st) {
  const size_t C = t->n_chains, d = (size_t)t->d;
  const size_t blk = t->block_chains ? t->block_chains : C;
  const size_t psz = (size_t)t->n_part * C;
  float* quad0 = t->part; float* quad1 = t->part + psz; float* ke1 = t->part + 2 * psz;
  const float eps = (float)S.eps, half = 0.5f * eps;
  int launches = 0;
  for (size_t c0 = 0; c0 < C; c0 += blk) {
    const size_t nb = std::min(blk, C - c0);
    const unsigned wblocks = (unsigned)((nb * 32 + 255) / 256);
    BeginArgs b;
    b.n_chains = nb; b.d = t->d; b.kpad = t->kpad; b.chain_offset = S.chain_offset + c0;
    b.key = PhiloxKey{(uint32_t)S.seed, (uint32_t)(S.seed >> 32)}; b.step = S.step;
    b.q = (const float*)S.q + c0 * d; b.p = t->p + c0 * d; b.mu = t->mu; b.dl = t->dl[0] + c0 * (size_t)t->kpad; b.ke0 = t->ke0 + c0;
    b.inj_normals = S.inj_normals ? (const float*)S.inj_normals + c0 * d : nullptr;
    b.dmax_bits = t->dmax_bits;
    dense_begin_kernel<<<wblocks, 256, 0, st>>>(b);
    dense_scale_kernel<<<1, 1, 0, st>>>(t->dmax_bits, t->dscale);
    launches += 2;
    // GEMM 0: gradient at the current point (log density, first half kick, drift of leapfrog 1);
    // GEMM l (1 <= l < L): kick eps + drift of leapfrog l + 1;  GEMM L: last half kick, log density, kinetic energy
    int final_buf = 0;
    if (S.n_leapfrog == 0) {
      if (gemm_kick(t, c0, nb, 0, 0.f, 0.f, quad0, ke1, st) != cudaSuccess) return -1;
      ++launches;
    } else {
      if (gemm_kick(t, c0, nb, 0, half, eps, quad0, nullptr, st) != cudaSuccess) return -1;
      ++launches;
      for (uint32_t l = 1; l <= S.n_leapfrog; ++l) {
        const bool last = (l == S.n_leapfrog);
        if (gemm_kick(t, c0, nb, (int)(l & 1u), last ? half : eps, last ? 0.f : eps, last ? quad1 : nullptr, last ? ke1 : nullptr, st) != cudaSuccess)
          return -1;
        ++launches;
      }
      final_buf = (int)(S.n_leapfrog & 1u);
    }
    AcceptArgs a;
    a.n_chains = nb; a.d = t->d; a.kpad = t->kpad; a.chain_offset = S.chain_offset + c0; a.key = b.key; a.step = S.step;
    a.q = (float*)S.q + c0 * d; a.dl = t->dl[final_buf] + c0 * (size_t)t->kpad; a.mu = t->mu; a.p = t->p + c0 * d;
    a.quad0 = quad0 + c0; a.quad1 = (S.n_leapfrog > 0 ? quad1 : quad0) + c0; a.ke0 = t->ke0 + c0; a.ke1_part = ke1 + c0;
    a.n_part = t->n_part; a.part_ld = C; a.norm_const = t->norm_const;
    a.out = S.out ? (float*)S.out + c0 * S.out_n * d : nullptr; a.out_n = S.out_n; a.slot = S.slot;
    a.accept_total = S.accept_total; a.diverge_total = S.diverge_total;
    a.inj_lnu = S.inj_lnu ? (const float*)S.inj_lnu + c0 : nullptr;
    a.diag_logacc = S.diag_logacc ? (float*)S.diag_logacc + c0 : nullptr; a.diag_acc = S.diag_acc ? S.diag_acc + c0 : nullptr;
    a.diag_pq = S.diag_pq ? (float*)S.diag_pq + c0 * d : nullptr; a.diag_pp = S.diag_pp ? (float*)S.diag_pp + c0 * d : nullptr;
    dense_accept_kernel<<<wblocks, 256, 0, st>>>(a);
    ++launches;
  }
  if (cudaGetLastError() != cudaSuccess) return -1;
  return launches;
}

}  // namespace gm
