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
stats_finalize(const float* __restrict__ spec, const double* __restrict__ mom, double total_chains /*unsplit, all ranks*/,
               size_t n, int p, int N, int log2n, const float2* __restrict__ tw, float* __restrict__ rhat,
               float* __restrict__ rhat_std, float* __restrict__ ess, float* __restrict__ acov_out /*[p][half] or null*/) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* re = reinterpret_cast<float*>(smem_raw);
  float* im = re + N;
  __shared__ float s_within, s_var;
  const int k = blockIdx.x;
  const int nk = N / 2 + 1;
  const size_t half = n / 2;
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

size_t stats_npad(size_t n) {
  const size_t half = n / 2;
  size_t N = 1;
  while (N < 2 * half - 1) N <<= 1;
  if (N < 2) N = 2;
  return N;
}

int stats_ppb(size_t N) {
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

cudaError_t launch_stats_finalize(const StatsLaunch& S, double total_chains, cudaStream_t st) {
  const size_t smem = S.N * 8;
  auto kern = stats_finalize;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<(unsigned)S.p, kStatsBlock, smem, st>>>(S.spec, S.mom, total_chains, S.n, S.p, (int)S.N, S.log2n,
                                                 (const float2*)S.tw, S.rhat, S.rhat_std, S.ess, S.acov);
  return cudaGetLastError();
}

}  // namespace gm
