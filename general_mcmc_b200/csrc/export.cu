// export.cu — device-side column export of a sample tensor (sm_100a).
//
// The reference's writers (/root/reference/src/io/csv.rs:47-147, io/arrow.rs:53-117, io/parquet.rs:49-222) walk the
// [chains, samples, dim] tensor on the host and append value by value to one builder per column: chain:u32,
// observation:u32, dim_0 .. dim_{d-1}:f64, one row per (chain, observation) pair — chain-major for save_csv / save_arrow /
// save_parquet, observation-major for save_parquet_tensor ([obs, chain, dim] order, parquet.rs:166-176).  Here the
// columns are produced on the device from the sample tensor where the sampler left it: a tiled transpose [rows, d] ->
// [d][rows] with the f32 -> f64 widening fused in (row-contiguous reads, column-contiguous writes), so a multi-GB tensor
// is never transposed element by element on the host.
#include "kernels.h"

namespace gm {

namespace {

constexpr int kTile = 32;

// grid (row tiles, dim tiles); block 32 x 8.  Output row r (0 .. C n - 1): chain-major r = c n + t, observation-major
// r = t C + c; the input row is always (c n + t).
// Columns [k0, k0 + nk) of the d-wide rows are exported (the host walks wide tensors in slabs of whole columns).
template <class TIN>
__global__ void __launch_bounds__(256) export_columns_kernel(const TIN* __restrict__ samples, size_t C, size_t n, int ld, int kbase, int d,
                                                             int obs_major, unsigned int chain_base,
                                                             unsigned int* __restrict__ chain_col, unsigned int* __restrict__ obs_col,
                                                             double* __restrict__ dims /*[d][rows]*/) {
  __shared__ float tile[kTile][kTile + 1];
  __shared__ double tile_d[sizeof(TIN) == 8 ? kTile : 1][sizeof(TIN) == 8 ? kTile + 1 : 1];
  const size_t rows = C * n;
  const size_t r0 = (size_t)blockIdx.x * kTile;
  const int k0 = blockIdx.y * kTile;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < kTile; i += 8) {
    const size_t r = r0 + i;
    if (r < rows && k0 + tx < d) {
      size_t c, t;
      if (obs_major) { t = r / C; c = r - t * C; } else { c = r / n; t = r - c * n; }
      const TIN v = samples[(c * n + t) * (size_t)ld + kbase + k0 + tx];
      if constexpr (sizeof(TIN) == 8) tile_d[i][tx] = v; else tile[i][tx] = (float)v;
    }
  }
  __syncthreads();
  for (int i = ty; i < kTile; i += 8) {
    const int k = k0 + i;
    const size_t r = r0 + tx;
    if (k < d && r < rows) {
      double v;
      if constexpr (sizeof(TIN) == 8) v = tile_d[tx][i]; else v = (double)tile[tx][i];
      dims[(size_t)k * rows + r] = v;
    }
  }
  if (blockIdx.y == 0 && ty == 0) {
    const size_t r = r0 + tx;
    if (r < rows) {
      size_t c, t;
      if (obs_major) { t = r / C; c = r - t * C; } else { c = r / n; t = r - c * n; }
      if (chain_col) chain_col[r] = chain_base + (unsigned int)c;
      if (obs_col) obs_col[r] = (unsigned int)t;
    }
  }
}

}  // namespace

cudaError_t launch_export_columns(const void* samples, int dtype, size_t C, size_t n, int ld, int kbase, int d, int obs_major,
                                  unsigned int chain_base, unsigned int* chain_col, unsigned int* obs_col, double* dims, cudaStream_t st) {
  const size_t rows = C * n;
  if (rows == 0) return cudaSuccess;
  const dim3 grid((unsigned)((rows + kTile - 1) / kTile), (unsigned)((d + kTile - 1) / kTile > 0 ? (d + kTile - 1) / kTile : 1));
  const dim3 block(kTile, 8);
  if (dtype == 0)
    export_columns_kernel<float><<<grid, block, 0, st>>>((const float*)samples, C, n, ld, kbase, d, obs_major, chain_base, chain_col, obs_col, dims);
  else
    export_columns_kernel<double><<<grid, block, 0, st>>>((const double*)samples, C, n, ld, kbase, d, obs_major, chain_base, chain_col, obs_col, dims);
  return cudaGetLastError();
}

}  // namespace gm
