// stats.cu — K4: device-side split R-hat / ESS reduction (sm_100a).
//
// Replaces the host loops of /root/reference/src/stats.rs:
//   split_rhat_mean_ess :439-450   splitcat :419-425   withinvar :456-504   rhat :452-454
//   ess :523-573 (Geyer initial-monotone on Stan's rho-hat)   autocov_fft :603-647 / autocov_bf :659-681
//
// The reference runs one FFT pair per (split chain, parameter) on the host and averages the
// autocovariances over chains.  Here the mean-over-chains autocovariance is obtained from the
// chain-SUMMED power spectrum (the inverse transform is linear), so the per-chain work is one forward
// FFT and a |Z|^2 accumulate, and the inverse FFT runs once per parameter:
//
//   stats_accumulate   grid (chain groups, parameter blocks of <=8).  Per chain: the [n, 8] tile is
//                      read as whole 32-byte sectors, the two halves of the chain (splitcat) are centred
//                      and packed as ONE complex series z = first + i*second; its power spectrum gives
//                      |X_k|^2 + |Y_k|^2 = (|Z_k|^2 + |Z_{N-k}|^2) / 2.  Per-chain means / within
//                      variances are accumulated in f64.  Per-group partials go to HBM (deterministic:
//                      no floating-point atomics).
//   stats_reduce       fixed-order sum of the group partials -> spectrum [p, N/2+1] f32, moments [p,3] f64.
//                      (a distributed context all-reduces these two buffers: collectives A2 + A3)
//   stats_finalize     one CTA per parameter: inverse FFT, W, B, var-hat, rho-hat, Geyer truncation.
//
// Zero-padding to N >= 2*half-1 makes the FFT autocovariance identical (up to f32 rounding) to the
// brute-force sum the reference uses for series of <= 100 draws, so one path serves both.
#include "kernels.h"

#include <cmath>
#include <cstdio>

namespace gm {

namespace {

constexpr int kStatsBlock = 512;
constexpr int kMaxPpb = 32;  // parameters per CTA (32 f32 = one 128-byte line per draw; fewer when the FFT is long)

__device__ __forceinline__ unsigned bitrev(unsigned v, int log2n) { return __brev(v) >> (32 - log2n); }

// In-place radix-2 DIT FFT over `nser` interleaved series in shared memory, input already in
// bit-reversed order.  re/im: [nser][N].  tw: [N/2] (cos, -sin) pairs of exp(-2 pi i k / N).
__device__ __forceinline__ void smem_fft(float* re, float* im, int nser, int N, int log2n,
                                         const float2* __restrict__ tw) {
  const int half_n = N >> 1;
  const int total = nser * half_n;
  for (int s = 1; s <= log2n; ++s) {
    const int hl = 1 << (s - 1);        // half length of this stage's butterflies
    const int tw_stride = N >> s;       // twiddle index stride
    for (int b = threadIdx.x; b < total; b += blockDim.x) {
      const int ser = b >> (log2n - 1);          // half_n = 2^(log2n - 1)
      const int bb = b & (half_n - 1);
      const int grp = bb >> (s - 1);
      const int k = bb & (hl - 1);
      const int i = ser * N + (grp << s) + k;
      const int j = i + hl;
      const float2 w = __ldg(tw + k * tw_stride);
      const float xr = re[j] * w.x - im[j] * w.y;
      const float xi = re[j] * w.y + im[j] * w.x;
      const float ar = re[i], ai = im[i];
      re[j] = ar - xr; im[j] = ai - xi;
      re[i] = ar + xr; im[i] = ai + xi;
    }
    __syncthreads();
  }
}

template <class TIN>
__global__ void __launch_bounds__(kStatsBlock)
stats_accumulate(const TIN* __restrict__ samples, size_t C, size_t n, int p, int ppb, int N, int log2n,
                 const float2* __restrict__ tw, float* __restrict__ part_spec /*[G][p][N/2+1]*/,
                 double* __restrict__ part_mom /*[G][p][3]*/) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* re = reinterpret_cast<float*>(smem_raw);          // [ppb][N]  first half of the chain
  float* im = re + (size_t)ppb * N;                        // [ppb][N]  second half
  float* acc = im + (size_t)ppb * N;                       // [ppb][N]  sum over chains of |Z_k|^2
  __shared__ double mom[kMaxPpb][3];

  const int G = gridDim.x;
  const int g = blockIdx.x;
  const int k0 = blockIdx.y * ppb;
  const int np = min(ppb, p - k0);                         // parameters handled by this CTA
  const size_t half = n / 2;
  const size_t off2 = n - half;                            // first draw of the second half
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = kStatsBlock / 32;

  for (int i = threadIdx.x; i < ppb * N; i += kStatsBlock) acc[i] = 0.f;
  if (threadIdx.x < kMaxPpb * 3) mom[threadIdx.x / 3][threadIdx.x % 3] = 0.0;   // kMaxPpb * 3 = 96 <= blockDim
  __syncthreads();

  // contiguous chain range of this group
  const size_t c_lo = (C * (size_t)g) / G, c_hi = (C * (size_t)(g + 1)) / G;
  for (size_t c = c_lo; c < c_hi; ++c) {
    // ---- A. tile [n, np] -> smem (bit-reversed slots), zero padding
    for (int i = threadIdx.x; i < ppb * N; i += kStatsBlock) { re[i] = 0.f; im[i] = 0.f; }
    __syncthreads();
    const TIN* base = samples + c * n * (size_t)p + k0;
    {
      const int j = threadIdx.x % ppb;
      const int rows_per_pass = kStatsBlock / ppb;
      if (j < np) {
        for (size_t t = threadIdx.x / ppb; t < half; t += rows_per_pass) {
          const unsigned slot = bitrev((unsigned)t, log2n);
          re[j * N + slot] = (float)base[t * p + j];
          im[j * N + slot] = (float)base[(off2 + t) * p + j];
        }
      }
    }
    __syncthreads();
    // ---- B. per (parameter, half): mean, centre in place, within variance (withinvar, stats.rs:456-504)
    for (int j = warp; j < np; j += nwarps) {   // one warp owns parameter j: deterministic f64 sums
      for (int h = 0; h < 2; ++h) {
        float* v = h ? (im + j * N) : (re + j * N);
        float s = 0.f;
        for (size_t t = lane; t < half; t += 32) s += v[bitrev((unsigned)t, log2n)];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)half;
        float sq = 0.f;
        for (size_t t = lane; t < half; t += 32) {
          const unsigned slot = bitrev((unsigned)t, log2n);
          const float cv = v[slot] - mean;
          v[slot] = cv;
          sq += cv * cv;
        }
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0) {
          mom[j][0] += (double)mean;
          mom[j][1] += (double)mean * (double)mean;
          mom[j][2] += (double)(sq / (float)half);
        }
      }
    }
    __syncthreads();
    // ---- C. forward FFT of z = first + i*second
    smem_fft(re, im, np, N, log2n, tw);
    // ---- D. accumulate |Z_k|^2
    for (int i = threadIdx.x; i < np * N; i += kStatsBlock) acc[i] += re[i] * re[i] + im[i] * im[i];
    __syncthreads();
  }

  // ---- partials: S_k = (A_k + A_{N-k}) / 2 for k = 0..N/2
  const int nk = N / 2 + 1;
  for (int i = threadIdx.x; i < np * nk; i += kStatsBlock) {
    const int j = i / nk, k = i - j * nk;
    const float a0 = acc[j * N + k], a1 = acc[j * N + ((N - k) & (N - 1))];
    part_spec[((size_t)g * p + k0 + j) * nk + k] = 0.5f * (a0 + a1);
  }
  if (threadIdx.x < np * 3) {
    const int j = threadIdx.x / 3, m = threadIdx.x % 3;
    part_mom[((size_t)g * p + k0 + j) * 3 + m] = mom[j][m];
  }
}

// ------------------------------------------------------------------------------------------------
// stats_accumulate_warp — the fast path of stats_accumulate for padded lengths 128 <= N <= 1024 (series of
// ~66 .. 1025 draws; BASELINE config 4 collects 500 or 1000).  Same partials, different organisation:
//   * a CTA of 8 warps owns 8 adjacent parameters (one 32-byte sector per draw, read exactly once) and a
//     contiguous range of chains; per chain the [n, 8] tile is loaded cooperatively, then EACH WARP transforms
//     one parameter's packed series on its own — no block-wide barrier inside the transform;
//   * the transform is a decimation-in-frequency FFT with radix-8 passes done in registers (one or two
//     butterflies per lane per pass; an optional final radix-2 / radix-4 pass), in place in shared memory
//     (re / im planes padded by one float per 8 so every pass is bank-conflict free), centring and the
//     within-variance sum fused into the first pass, |Z_k|^2 accumulated in registers straight out of the
//     last pass (digit-reversed order; the order is undone once, when the partials are written);
//   * twiddles: one table entry w_L^j per butterfly, its powers by complex multiplication.
// ------------------------------------------------------------------------------------------------
constexpr int kWarpFftWarps = 8;   // warps per CTA = parameters per CTA
// resident CTAs per SM of the warp-FFT kernel for N <= 512.  Measured at 65,536 x 500 x 100: 2 CTAs (125 registers, no spills)
// 10.8 ms, 3 CTAs (80 registers) 9.6 ms, 4 CTAs (64 registers, spills) 10.5 ms (profiles/r2_k4_occupancy_scan.txt)
#ifndef GM_STATS_MINB
#define GM_STATS_MINB 3
#endif

__device__ __forceinline__ int padi(int a) { return a + (a >> 3); }

__device__ __forceinline__ void cmul(float& xr, float& xi, float wr, float wi) {
  const float t = xr * wr - xi * wi;
  xi = xr * wi + xi * wr;
  xr = t;
}

// 4-point DFT, natural order in and out
__device__ __forceinline__ void dft4(float& r0, float& i0, float& r1, float& i1, float& r2, float& i2, float& r3, float& i3) {
  const float p0r = r0 + r2, p0i = i0 + i2, p2r = r0 - r2, p2i = i0 - i2;
  const float p1r = r1 + r3, p1i = i1 + i3;
  const float p3r = i1 - i3, p3i = r3 - r1;          // (a1 - a3) * (-i)
  r0 = p0r + p1r; i0 = p0i + p1i;
  r2 = p0r - p1r; i2 = p0i - p1i;
  r1 = p2r + p3r; i1 = p2i + p3i;
  r3 = p2r - p3r; i3 = p2i - p3i;
}

// 8-point DFT, natural order in and out: X[2m] = DFT4(x_i + x_{i+4})[m], X[2m+1] = DFT4((x_i - x_{i+4}) w8^i)[m]
__device__ __forceinline__ void dft8(float (&xr)[8], float (&xi)[8]) {
  constexpr float c = 0.70710678118654752440f;
  float ur[4], ui[4], vr[4], vi[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ur[i] = xr[i] + xr[i + 4]; ui[i] = xi[i] + xi[i + 4];
    vr[i] = xr[i] - xr[i + 4]; vi[i] = xi[i] - xi[i + 4];
  }
  { const float a = vr[1], b = vi[1]; vr[1] = (a + b) * c; vi[1] = (b - a) * c; }      // * w8
  { const float a = vr[2], b = vi[2]; vr[2] = b; vi[2] = -a; }                          // * (-i)
  { const float a = vr[3], b = vi[3]; vr[3] = (b - a) * c; vi[3] = -(a + b) * c; }     // * w8^3
  dft4(ur[0], ui[0], ur[1], ui[1], ur[2], ui[2], ur[3], ui[3]);
  dft4(vr[0], vi[0], vr[1], vi[1], vr[2], vi[2], vr[3], vi[3]);
#pragma unroll
  for (int m = 0; m < 4; ++m) { xr[2 * m] = ur[m]; xi[2 * m] = ui[m]; xr[2 * m + 1] = vr[m]; xi[2 * m + 1] = vi[m]; }
}

// the same with x[4 .. 7] == 0 (zero padding): u_i = v_i = x_i
__device__ __forceinline__ void dft8_upper_zero(float (&xr)[8], float (&xi)[8]) {
  constexpr float c = 0.70710678118654752440f;
  float ur[4], ui[4], vr[4], vi[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { ur[i] = xr[i]; ui[i] = xi[i]; vr[i] = xr[i]; vi[i] = xi[i]; }
  { const float a = vr[1], b = vi[1]; vr[1] = (a + b) * c; vi[1] = (b - a) * c; }      // * w8
  { const float a = vr[2], b = vi[2]; vr[2] = b; vi[2] = -a; }                          // * (-i)
  { const float a = vr[3], b = vi[3]; vr[3] = (b - a) * c; vi[3] = -(a + b) * c; }     // * w8^3
  dft4(ur[0], ui[0], ur[1], ui[1], ur[2], ui[2], ur[3], ui[3]);
  dft4(vr[0], vi[0], vr[1], vi[1], vr[2], vi[2], vr[3], vi[3]);
#pragma unroll
  for (int m = 0; m < 4; ++m) { xr[2 * m] = ur[m]; xi[2 * m] = ui[m]; xr[2 * m + 1] = vr[m]; xi[2 * m + 1] = vi[m]; }
}

template <int LOG2N>
struct WarpFft {
  static constexpr int N = 1 << LOG2N;
  static constexpr int NP8 = LOG2N / 3;                 // radix-8 passes
  static constexpr int RL = 1 << (LOG2N % 3);           // radix of the final short pass (1: none)
  // plane length: one pad float per 8 (the passes' strides) plus 2, so that a parameter's plane pair starts 4 banks after
  // its neighbour's: the tile loader's stores (8 parameters x 4 consecutive draws per warp instruction) then hit 32 different
  // banks — with N + N / 8 every plane started on bank 0 and those stores were 8-way conflicts.
  static constexpr int PADN = N + N / 8 + 2;
  static constexpr int LASTR = RL > 1 ? RL : 8;
  static constexpr int NB_LAST = (N / LASTR + 31) / 32;
  static constexpr int ACC = LASTR * NB_LAST;           // |Z|^2 accumulators per lane
  static constexpr int TW = N / 7 + 8;                  // twiddle table entries (sum of M over the radix-8 passes with M > 1)

  // position in the transformed array -> frequency index (digit reversal of the mixed-radix schedule)
  __device__ static int freq_of(int pos) {
    int k = 0, mult = 1, rem = pos, span = N;
#pragma unroll
    for (int ps = 0; ps < NP8; ++ps) { span >>= 3; const int q = rem / span; rem -= q * span; k += q * mult; mult <<= 3; }
    if (RL > 1) k += rem * mult;
    return k;
  }

  // radix-8 pass PS.  FIRST: centre (subtract the half's mean), zero-pad beyond `half`, accumulate the squared deviations.
  // LAST: accumulate |y|^2 instead of storing.
  template <int PS, bool FIRST, bool LAST>
  __device__ static __forceinline__ void pass8(float* re, float* im, const float2* tw, int lane, int half, float mean0,
                                               float mean1, float& sq0, float& sq1, float (&acc)[ACC]) {
    constexpr int L = N >> (3 * PS), M = L >> 3, NBF = N / 8, NB = (NBF + 31) / 32;
    // the last pass indexes the register accumulators: fully unrolled; the others two butterflies at a time
#pragma unroll(LAST ? NB : (NB < 2 ? NB : 2))
    for (int i = 0; i < NB; ++i) {
      const int t = lane + 32 * i;
      if (NBF % 32 == 0 || t < NBF) {
        const int b = t / M, j = t - b * M;
        const int base = b * L + j;
        // padded address of element base + r M: M is a multiple of 8 (constant padded stride) or 1 (base is a multiple of 8)
        // (spans 2 and 4 of the mixed-radix lengths fall back to the per-element form)
        constexpr bool kLin = (M % 8 == 0) || M == 1;
        constexpr int PSTR = (M % 8 == 0) ? M + M / 8 : M;
        const int pb = padi(base);
        auto at = [&](int r) { return kLin ? pb + r * PSTR : padi(base + r * M); };
        float xr[8], xi[8];
        if (FIRST) {
          // inputs r = 4 .. 7 sit at indices >= N / 2 >= half: the zero padding.  Only four loads, and the first butterfly
          // stage of the 8-point DFT (x_i +- x_{i+4}) degenerates to copies.
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const bool valid = base + r * M < half;
            const float vr = valid ? re[at(r)] - mean0 : 0.f;
            const float vi = valid ? im[at(r)] - mean1 : 0.f;
            sq0 += vr * vr; sq1 += vi * vi;
            xr[r] = vr; xi[r] = vi;
          }
          dft8_upper_zero(xr, xi);
        } else {
#pragma unroll
          for (int r = 0; r < 8; ++r) { xr[r] = re[at(r)]; xi[r] = im[at(r)]; }
          dft8(xr, xi);
        }
        if (M > 1) {
          const float2 w1 = tw[j];
          float wr = w1.x, wi = w1.y;
          cmul(xr[1], xi[1], wr, wi);
          float w2r = wr, w2i = wi; cmul(w2r, w2i, wr, wi);
          cmul(xr[2], xi[2], w2r, w2i);
          float w3r = w2r, w3i = w2i; cmul(w3r, w3i, wr, wi);
          cmul(xr[3], xi[3], w3r, w3i);
          float w4r = w2r, w4i = w2i; cmul(w4r, w4i, w2r, w2i);
          cmul(xr[4], xi[4], w4r, w4i);
          float w5r = w4r, w5i = w4i; cmul(w5r, w5i, wr, wi);
          cmul(xr[5], xi[5], w5r, w5i);
          float w6r = w4r, w6i = w4i; cmul(w6r, w6i, w2r, w2i);
          cmul(xr[6], xi[6], w6r, w6i);
          float w7r = w4r, w7i = w4i; cmul(w7r, w7i, w3r, w3i);
          cmul(xr[7], xi[7], w7r, w7i);
        }
        if (LAST) {
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[i * 8 + q] += xr[q] * xr[q] + xi[q] * xi[q];
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) { re[at(q)] = xr[q]; im[at(q)] = xi[q]; }
        }
      }
    }
  }

  // final radix-2 / radix-4 pass (span RL, no twiddles), accumulating |y|^2
  __device__ static __forceinline__ void pass_last_short(const float* re, const float* im, int lane, float (&acc)[ACC]) {
    constexpr int NBF = N / RL;
#pragma unroll
    for (int i = 0; i < NB_LAST; ++i) {
      const int t = lane + 32 * i;
      if (NBF % 32 == 0 || t < NBF) {
        const int base = padi(t * RL);          // RL consecutive elements never straddle a pad slot
        if (RL == 2) {
          const float ar = re[base], ai = im[base], br = re[base + 1], bi = im[base + 1];
          const float sr = ar + br, si = ai + bi, dr = ar - br, di = ai - bi;
          acc[i * LASTR + 0] += sr * sr + si * si;
          acc[i * LASTR + 1] += dr * dr + di * di;
        } else if (RL == 4) {
          float r0 = re[base], i0 = im[base], r1 = re[base + 1], i1 = im[base + 1];
          float r2 = re[base + 2], i2 = im[base + 2], r3 = re[base + 3], i3 = im[base + 3];
          dft4(r0, i0, r1, i1, r2, i2, r3, i3);
          acc[i * LASTR + 0] += r0 * r0 + i0 * i0;
          acc[i * LASTR + 1] += r1 * r1 + i1 * i1;
          acc[i * LASTR + 2] += r2 * r2 + i2 * i2;
          acc[i * LASTR + 3] += r3 * r3 + i3 * i3;
        }
      }
    }
  }

  template <int PS>
  __device__ static __forceinline__ void passes(float* re, float* im, const float2* tw, int lane, int half, float mean0,
                                                float mean1, float& sq0, float& sq1, float (&acc)[ACC]) {
    if constexpr (PS < NP8) {
      constexpr bool last = (PS == NP8 - 1) && RL == 1;
      pass8<PS, PS == 0, last>(re, im, tw, lane, half, mean0, mean1, sq0, sq1, acc);
      if constexpr (!last) __syncwarp();
      constexpr int M = (N >> (3 * PS)) >> 3;
      passes<PS + 1>(re, im, tw + (M > 1 ? M : 0), lane, half, mean0, mean1, sq0, sq1, acc);
    } else if constexpr (RL > 1) {
      pass_last_short(re, im, lane, acc);
    }
  }
};

template <class TIN, int LOG2N>
__global__ void __launch_bounds__(kWarpFftWarps * 32, (LOG2N >= 10 ? 2 : GM_STATS_MINB))
stats_accumulate_warp(const TIN* __restrict__ samples, size_t C, size_t n, int p, float* __restrict__ part_spec /*[G][p][N/2+1]*/,
                      double* __restrict__ part_mom /*[G][p][3]*/) {
  using F = WarpFft<LOG2N>;
  constexpr int N = F::N, PADN = F::PADN;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* planes = reinterpret_cast<float*>(smem_raw);                 // [8 warps][re, im][PADN]
  __shared__ float2 tw[F::TW];

  // 1-D grid, parameter block fastest: the CTAs that read neighbouring 32-byte sectors of the same draws are
  // co-scheduled, so the sectors DRAM fetches alongside are L2 hits instead of being fetched again later
  const int pblocks = (p + kWarpFftWarps - 1) / kWarpFftWarps;
  const int G = gridDim.x / pblocks, g = blockIdx.x / pblocks;
  const int k0 = (blockIdx.x - g * pblocks) * kWarpFftWarps;
  const int np = min(kWarpFftWarps, p - k0);
  // (double-buffering the planes to drop one of the two block barriers per chain was measured slower: 17.4 vs 13.3 ms)
  __shared__ float part_sum[kWarpFftWarps][2][kWarpFftWarps];   // [parameter][half][loading warp]
  const int half = (int)(n / 2);
  const size_t off2 = n - (size_t)half;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* re = planes + (size_t)warp * 2 * PADN;
  float* im = re + PADN;

  {  // twiddle tables of the radix-8 passes with M > 1: w_L^j = exp(-2 pi i j / L), j < M = L / 8
    int off = 0;
#pragma unroll
    for (int ps = 0; ps < F::NP8; ++ps) {
      const int L = N >> (3 * ps), M = L >> 3;
      if (M > 1) {
        for (int j = threadIdx.x; j < M; j += blockDim.x) {
          float sn, cs;
          sincospif(-2.0f * (float)j / (float)L, &sn, &cs);
          tw[off + j] = make_float2(cs, sn);
        }
        off += M;
      }
    }
  }
  float acc[F::ACC];
#pragma unroll
  for (int i = 0; i < F::ACC; ++i) acc[i] = 0.f;
  double m0 = 0.0, m1 = 0.0, m2 = 0.0;
  const float inv_half = 1.0f / (float)half;

  // ---- A. the chain's [n, np] tile -> the warps' planes: re = first half, im = second half (splitcat).
  // Thread (warp w, lane l) loads parameter j = l & 7 of draws t = 4 w + (l >> 3) + 32 m.  The loads of chain c + 1 are
  // issued into registers BEFORE the transform of chain c and committed to shared memory after it, so their latency
  // hides under the butterflies; the loaders' running sums feed the means.
  constexpr int NLD = (N / 2 + 31) / 32;
  const int lj = lane & 7, lt0 = threadIdx.x >> 3;
  float pa[NLD], pb[NLD];
  // addresses: one uniform base per (chain, half) and a 32-bit element offset per thread that advances by 32 draws per
  // load (a chain's [n, p] block is far below 2^32 elements) — the 64-bit multiply per load of the first version was a
  // fifth of the kernel's instructions
  const unsigned row_step = 32u * (unsigned)p;
  const unsigned off0 = (unsigned)lt0 * (unsigned)p + (unsigned)lj;
  const bool lane_on = lj < np;
  auto prefetch = [&](size_t c) {
    const TIN* base0 = samples + c * n * (size_t)p + k0;
    const TIN* base1 = base0 + off2 * (size_t)p;
    unsigned off = off0;
#pragma unroll
    for (int m = 0; m < NLD; ++m) {
      const bool on = lane_on && (lt0 + 32 * m) < half;
      pa[m] = on ? (float)base0[off] : 0.f;
      pb[m] = on ? (float)base1[off] : 0.f;
      off += row_step;
    }
  };
  auto commit = [&]() {
    float* pr = planes + (size_t)lj * 2 * PADN;
    float ls0 = 0.f, ls1 = 0.f;
    // draws at or beyond `half` are stored too (as the zeros prefetch left): t <= 32 NLD - 1 < N, inside the plane, and the
    // first pass masks them anyway — no branch per store
#pragma unroll
    for (int m = 0; m < NLD; ++m) {
      const int pt = padi(lt0 + 32 * m);
      pr[pt] = pa[m]; pr[PADN + pt] = pb[m];
      ls0 += pa[m]; ls1 += pb[m];
    }
    ls0 += __shfl_xor_sync(0xffffffffu, ls0, 8); ls1 += __shfl_xor_sync(0xffffffffu, ls1, 8);
    ls0 += __shfl_xor_sync(0xffffffffu, ls0, 16); ls1 += __shfl_xor_sync(0xffffffffu, ls1, 16);
    if (lane < 8) { part_sum[lane][0][warp] = ls0; part_sum[lane][1][warp] = ls1; }
  };

  const size_t c_lo = (C * (size_t)g) / G, c_hi = (C * (size_t)(g + 1)) / G;
  if (c_lo < c_hi) prefetch(c_lo);
  for (size_t c = c_lo; c < c_hi; ++c) {
    commit();
    __syncthreads();
    if (c + 1 < c_hi) prefetch(c + 1);
    if (warp < np) {
      // ---- B. means of the two halves (withinvar, stats.rs:456-504): fixed-order sum of the loaders' partials
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int w = 0; w < kWarpFftWarps; ++w) { s0 += part_sum[warp][0][w]; s1 += part_sum[warp][1][w]; }
      const float mean0 = s0 * inv_half, mean1 = s1 * inv_half;
      // ---- C. forward FFT of z = (first - mean0) + i (second - mean1), |Z|^2 accumulated out of the last pass
      float sq0 = 0.f, sq1 = 0.f;
      F::template passes<0>(re, im, tw, lane, half, mean0, mean1, sq0, sq1, acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { sq0 += __shfl_xor_sync(0xffffffffu, sq0, o); sq1 += __shfl_xor_sync(0xffffffffu, sq1, o); }
      m0 += (double)mean0; m1 += (double)mean0 * (double)mean0; m2 += (double)(sq0 * inv_half);
      m0 += (double)mean1; m1 += (double)mean1 * (double)mean1; m2 += (double)(sq1 * inv_half);
    }
    __syncthreads();
  }

  // ---- partials: A_k back in frequency order (through this warp's re plane), S_k = (A_k + A_{N-k}) / 2 for k = 0..N/2
  if (warp < np) {
#pragma unroll
    for (int i = 0; i < F::NB_LAST; ++i) {
      const int t = lane + 32 * i;
      if ((N / F::LASTR) % 32 == 0 || t < N / F::LASTR) {
#pragma unroll
        for (int q = 0; q < F::LASTR; ++q) re[F::freq_of(t * F::LASTR + q)] = acc[i * F::LASTR + q];
      }
    }
    __syncwarp();
    constexpr int nk = N / 2 + 1;
    float* dst = part_spec + ((size_t)g * p + k0 + warp) * nk;
    for (int k = lane; k < nk; k += 32) dst[k] = 0.5f * (re[k] + re[(N - k) & (N - 1)]);
    if (lane == 0) {
      double* md = part_mom + ((size_t)g * p + k0 + warp) * 3;
      md[0] = m0; md[1] = m1; md[2] = m2;
    }
  }
}

static size_t warp_fft_smem(size_t N) { return (size_t)kWarpFftWarps * 2 * (N + N / 8 + 2) * sizeof(float); }

template <class TIN>
cudaError_t launch_accumulate_warp(const StatsLaunch& S, cudaStream_t st) {
  const unsigned grid = (unsigned)S.n_groups * (unsigned)((S.p + kWarpFftWarps - 1) / kWarpFftWarps);
  const TIN* x = (const TIN*)S.samples;
#define GM_WARP_FFT_CASE(LG)                                                                                     \
  case LG: {                                                                                                     \
    auto kern = stats_accumulate_warp<TIN, LG>;                                                                  \
    const size_t smem = warp_fft_smem(WarpFft<LG>::N);                                                           \
    if (smem > 48 * 1024) {                                                                                      \
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
      if (e != cudaSuccess) return e;                                                                            \
    }                                                                                                            \
    kern<<<grid, kWarpFftWarps * 32, smem, st>>>(x, S.C, S.n, S.p, S.part_spec, S.part_mom);                     \
    break;                                                                                                       \
  }
  switch (S.log2n) {
    GM_WARP_FFT_CASE(7) GM_WARP_FFT_CASE(8) GM_WARP_FFT_CASE(9) GM_WARP_FFT_CASE(10)
    default: return cudaErrorInvalidValue;
  }
#undef GM_WARP_FFT_CASE
  return cudaGetLastError();
}

__global__ void stats_reduce(const float* __restrict__ part_spec, const double* __restrict__ part_mom, int G,
                             size_t spec_len /*p*nk*/, size_t mom_len /*p*3*/, float* __restrict__ spec,
                             double* __restrict__ mom) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < spec_len) {
    double s = 0.0;
    for (int g = 0; g < G; ++g) s += (double)part_spec[(size_t)g * spec_len + i];
    spec[i] = (float)s;
  }
  if (i < mom_len) {
    double s = 0.0;
    for (int g = 0; g < G; ++g) s += part_mom[(size_t)g * mom_len + i];
    mom[i] = s;
  }
}

// one CTA per parameter
__global__ void __launch_bounds__(kStatsBlock)
stats_finalize(const float* __restrict__ spec, const double* __restrict__ mom, double total_chains_host /*unsplit, all ranks*/,
               const double* __restrict__ total_chains_dev /*same, device-resident (after an all-reduce), or null*/,
               size_t n, int p, int N, int log2n, const float2* __restrict__ tw, float* __restrict__ rhat,
               float* __restrict__ rhat_std, float* __restrict__ ess, float* __restrict__ acov_out /*[p][half] or null*/) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* re = reinterpret_cast<float*>(smem_raw);
  float* im = re + N;
  __shared__ float s_within, s_var;
  const int k = blockIdx.x;
  const int nk = N / 2 + 1;
  const size_t half = n / 2;
  const double total_chains = total_chains_dev ? *total_chains_dev : total_chains_host;
  const double c2 = 2.0 * total_chains;

  // symmetric real spectrum, bit-reversed slots; FFT of a real even sequence == N * inverse FFT
  for (int i = threadIdx.x; i < N; i += kStatsBlock) {
    const int kk = (i <= N / 2) ? i : (N - i);
    const unsigned slot = (log2n > 0) ? bitrev((unsigned)i, log2n) : 0u;
    re[slot] = spec[(size_t)k * nk + kk];
    im[slot] = 0.f;
  }
  if (threadIdx.x == 0) {
    // withinvar, stats.rs:456-504 with split shape (c2 chains, `half` draws); sums carried in f64
    const double sm = mom[(size_t)k * 3 + 0], sm2 = mom[(size_t)k * 3 + 1], sw = mom[(size_t)k * 3 + 2];
    const double om = sm / c2;
    double ss = sm2 - c2 * om * om;
    if (ss < 0.0) ss = 0.0;
    const double b = ss * ((double)half / (c2 - 1.0));
    const double w = sw / c2;
    const double v = (((double)half - 1.0) / (double)half) * w + b / (double)half;
    s_within = (float)w;
    s_var = (float)v;
  }
  __syncthreads();
  smem_fft(re, im, 1, N, log2n, tw);
  // mean autocovariance: re[t] / N / half / c2   (autocov_fft normalisation stats.rs:640-645, then mean over chains)
  const float within = s_within, var = s_var;
  for (size_t t = threadIdx.x; t < half; t += kStatsBlock) {
    const float ac = (float)((double)re[t] / (double)N / (double)half / c2);
    if (acov_out) acov_out[(size_t)k * half + t] = ac;
    const float diff = -ac + within;          // stats.rs:536-539
    re[t] = -(diff / var) + 1.0f;             // rho-hat, stats.rs:540-544
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // Geyer initial monotone sequence, stats.rs:545-570
    float mn = half >= 2 ? re[0] + re[1] : 0.0f;
    float out = 0.0f;
    for (size_t t = 0; t + 1 < half; t += 2) {
      float p_t = re[t] + re[t + 1];
      if (p_t <= 0.0f) break;
      if (p_t > mn) p_t = mn;
      mn = p_t;
      out += p_t;
    }
    const float tau = -1.0f + 2.0f * out;
    ess[k] = (1.0f / tau) * (float)c2 * (float)half;   // stats.rs:572
    rhat[k] = sqrtf(within / var);                      // stats.rs:452-454 (reference orientation)
    rhat_std[k] = sqrtf(var / within);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// K6 — progress tracker (MultiChainTracker, stats.rs:199-339) evaluated on the device from a [C, n, p] tensor of
// draws: what the reference accumulates on the host after every step of run_progress (running mean / mean of
// squares per chain and parameter, EMA acceptance rate, tracker R-hat) computed in one pass over the tensor.
// ------------------------------------------------------------------------------------------------
// One thread per (chain, parameter): the reference's f32 recurrences in the reference's order, no contraction
// (stats.rs:249-255):  mean = (mean (n-1) + x) / n,  mean_sq = x^2 for n = 1, else (mean_sq (n-1) + x^2) / n.
template <class TIN>
__global__ void __launch_bounds__(256) tracker_moments_kernel(const TIN* __restrict__ x, size_t C, size_t n, int p,
                                                              float* __restrict__ mean, float* __restrict__ mean_sq) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * (size_t)p) return;
  const size_t c = i / (size_t)p;
  const TIN* src = x + c * n * (size_t)p + (i - c * (size_t)p);
  float m = 0.f, q = 0.f;
#pragma unroll 8
  for (size_t t = 0; t < n; ++t) {
    const float v = (float)src[t * (size_t)p];
    const float nf = (float)(t + 1), nm1 = __fadd_rn(nf, -1.0f);
    const float vv = __fmul_rn(v, v);
    m = __fdiv_rn(__fadd_rn(__fmul_rn(m, nm1), v), nf);
    q = (t == 0) ? vv : __fdiv_rn(__fadd_rn(__fmul_rn(q, nm1), vv), nf);
  }
  mean[i] = m; mean_sq[i] = q;
}

// One CTA per parameter: within / between / var-hat / R-hat of MultiChainTracker::within_and_var (stats.rs:314-339);
// the sums over chains are carried in f64 in a fixed order.
__global__ void __launch_bounds__(256) tracker_rhat_kernel(const float* __restrict__ mean, const float* __restrict__ mean_sq,
                                                           size_t C, size_t n, int p, float* __restrict__ rhat) {
  __shared__ double red[256];
  const int k = blockIdx.x;
  auto block_sum = [&](double v) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    const double r = red[0];
    __syncthreads();
    return r;
  };
  double s = 0.0;
  for (size_t c = threadIdx.x; c < C; c += 256) s += (double)mean[c * p + k];
  const float nc = (float)C, nf = (float)n;
  const float mc = (float)(block_sum(s) / (double)C);
  double b = 0.0, w = 0.0;
  for (size_t c = threadIdx.x; c < C; c += 256) {
    const float m1 = mean[c * p + k];
    const float df = m1 - mc;
    b += (double)(df * df);
    w += (double)((mean_sq[c * p + k] - m1 * m1) * nf / (nf - 1.0f));
  }
  const double bs = block_sum(b), ws = block_sum(w);
  if (threadIdx.x == 0) {
    const float between = (float)bs * (nf / (nc - 1.0f));
    const float within = (float)(ws / (double)C);
    const float var = within * ((nf - 1.0f) / nf) + between * (1.0f / nf);
    rhat[k] = sqrtf(var / within);
  }
}

// EMA acceptance rate (stats.rs:257-265): p <- 0.99 p + 0.01 [row changed], folded over the chains of a step, step
// after step, from p = 0 and last_state = 0.  Terms older than 4096 folds carry a weight below 0.99^4096 = 1e-18, so
// folding the LAST 4096 (step, chain) pairs sequentially in f32 reproduces the full fold to the last bit.
template <class TIN>
__global__ void __launch_bounds__(1024) tracker_accept_kernel(const TIN* __restrict__ x, size_t C, size_t n, int p,
                                                              float* __restrict__ p_accept) {
  constexpr int W = 4096;
  __shared__ unsigned char flag[W];
  const size_t total = C * n;
  const size_t first = total > (size_t)W ? total - W : 0;
  const int cnt = (int)(total - first);
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const size_t kk = first + i;                 // fold index = t * C + c
    const size_t t = kk / C, c = kk - t * C;
    const TIN* cur = x + (c * n + t) * (size_t)p;
    bool ne = false;
    for (int j = 0; j < p; ++j) {
      const float a = (float)cur[j];
      const float b = t > 0 ? (float)cur[j - p] : 0.f;
      ne = ne || (a != b);
    }
    flag[i] = ne ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float pa = 0.f;
    for (int i = 0; i < cnt; ++i) pa = __fadd_rn(__fmul_rn(1.0f - 0.01f, pa), __fmul_rn(0.01f, flag[i] ? 1.0f : 0.0f));
    *p_accept = pa;
  }
}

cudaError_t launch_tracker(const void* samples, int dtype, size_t C, size_t n, int p, float* mean, float* mean_sq,
                           float* rhat, float* p_accept, cudaStream_t st) {
  const size_t total = C * (size_t)p;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (dtype == 0) {
    tracker_moments_kernel<float><<<blocks, 256, 0, st>>>((const float*)samples, C, n, p, mean, mean_sq);
    tracker_accept_kernel<float><<<1, 1024, 0, st>>>((const float*)samples, C, n, p, p_accept);
  } else {
    tracker_moments_kernel<double><<<blocks, 256, 0, st>>>((const double*)samples, C, n, p, mean, mean_sq);
    tracker_accept_kernel<double><<<1, 1024, 0, st>>>((const double*)samples, C, n, p, p_accept);
  }
  tracker_rhat_kernel<<<(unsigned)p, 256, 0, st>>>(mean, mean_sq, C, n, p, rhat);
  return cudaGetLastError();
}

size_t stats_npad(size_t n) {
  const size_t half = n / 2;
  size_t N = 1;
  while (N < 2 * half - 1) N <<= 1;
  if (N < 2) N = 2;
  return N;
}

int stats_ppb(size_t N);
bool stats_warp_path(size_t N) { return N >= 128 && N <= 1024; }

// chain groups (grid.x of the accumulate kernel): about two full waves of resident CTAs
int stats_groups(size_t N, size_t p, size_t C, int sm_count) {
  const int ppb = stats_ppb(N);
  const int pblocks = (int)((p + ppb - 1) / ppb);
  int target = sm_count * 4;
  if (stats_warp_path(N)) {
    int occ = 0, lg = 0;
    while (((size_t)1 << lg) < N) ++lg;
    const size_t smem = warp_fft_smem(N);
    cudaError_t e = cudaErrorInvalidValue;
#define GM_OCC_CASE(LG)                                                                                  \
  case LG: {                                                                                             \
    auto kern = stats_accumulate_warp<float, LG>;                                                        \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarpFftWarps * 32, smem);             \
    break;                                                                                               \
  }
    switch (lg) { GM_OCC_CASE(7) GM_OCC_CASE(8) GM_OCC_CASE(9) GM_OCC_CASE(10) }
#undef GM_OCC_CASE
    if (e != cudaSuccess || occ < 1) { occ = 1; (void)cudaGetLastError(); }
    target = sm_count * occ * 2;
  }
  int groups = (target + pblocks - 1) / pblocks;
  if ((size_t)groups > C) groups = (int)C;
  return groups < 1 ? 1 : groups;
}

int stats_ppb(size_t N) {
  if (stats_warp_path(N)) return kWarpFftWarps;   // stats_accumulate_warp: one warp per parameter, 8 per CTA
  // re + im + acc = 12 bytes per (parameter, bin); keep the CTA under ~192 KB of shared memory
  int ppb = kMaxPpb;
  while (ppb > 1 && (size_t)ppb * N * 12 > 200 * 1024) ppb >>= 1;
  return ppb;
}

void stats_fill_twiddles(size_t N, float* host_tw /*[N/2][2]*/) {
  for (size_t k = 0; k < N / 2; ++k) {
    const double ang = -2.0 * 3.14159265358979323846264338327950288 * (double)k / (double)N;
    host_tw[2 * k] = (float)std::cos(ang);
    host_tw[2 * k + 1] = (float)std::sin(ang);
  }
}

cudaError_t launch_stats_accumulate(const StatsLaunch& S, cudaStream_t st) {
  if (stats_warp_path(S.N) && S.ppb == kWarpFftWarps)
    return S.dtype == 0 ? launch_accumulate_warp<float>(S, st) : launch_accumulate_warp<double>(S, st);
  const int ppb = S.ppb;
  const size_t smem = (size_t)ppb * S.N * 12;
  dim3 grid((unsigned)S.n_groups, (unsigned)((S.p + ppb - 1) / ppb));
  if (S.dtype == 0) {
    auto kern = stats_accumulate<float>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kStatsBlock, smem, st>>>((const float*)S.samples, S.C, S.n, S.p, ppb, (int)S.N, S.log2n,
                                          (const float2*)S.tw, S.part_spec, S.part_mom);
  } else {
    auto kern = stats_accumulate<double>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kStatsBlock, smem, st>>>((const double*)S.samples, S.C, S.n, S.p, ppb, (int)S.N, S.log2n,
                                          (const float2*)S.tw, S.part_spec, S.part_mom);
  }
  return cudaGetLastError();
}

cudaError_t launch_stats_reduce(const StatsLaunch& S, cudaStream_t st) {
  const size_t nk = S.N / 2 + 1;
  const size_t spec_len = (size_t)S.p * nk, mom_len = (size_t)S.p * 3;
  const size_t len = spec_len > mom_len ? spec_len : mom_len;
  const unsigned blocks = (unsigned)((len + 255) / 256);
  stats_reduce<<<blocks, 256, 0, st>>>(S.part_spec, S.part_mom, S.n_groups, spec_len, mom_len, S.spec, S.mom);
  return cudaGetLastError();
}

cudaError_t launch_stats_finalize(const StatsLaunch& S, double total_chains, const double* total_chains_dev, cudaStream_t st) {
  const size_t smem = S.N * 8;
  auto kern = stats_finalize;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<(unsigned)S.p, kStatsBlock, smem, st>>>(S.spec, S.mom, total_chains, total_chains_dev, S.n, S.p, (int)S.N, S.log2n,
                                                 (const float2*)S.tw, S.rhat, S.rhat_std, S.ess, S.acov);
  return cudaGetLastError();
}

}  // namespace gm
