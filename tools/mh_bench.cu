// tools/mh_bench.cu — stand-alone timing harness for K2 (the 2-D fast MH kernel) used to compare compile-time
// variants without rebuilding libgmcmc: nvcc ... -DGM_MH2_MINB=7 tools/mh_bench.cu -o tools/mh_bench_m7
// Workload = BASELINE config 2 (Gaussian2D, identity covariance, 1,048,576 chains x 1000 steps, f64).
#include "../general_mcmc_b200/csrc/mh_kernel.cuh"
#include <cstdio>
#include <vector>
int main(int argc, char** argv) {
  const size_t C = 1048576; const uint32_t n = 1000;
  double* state; double* out; unsigned long long* acc;
  cudaMalloc(&state, C * 2 * 8); cudaMalloc(&out, C * n * 2 * 8); cudaMalloc(&acc, 8);
  cudaMemset(state, 0, C * 16); cudaMemset(acc, 0, 8);
  gm::MhLaunch L{};
  L.tgt.kind = 1; L.tgt.dtype = 1; L.tgt.dim = 2;
  double sp[8] = {0, 0, 1, 0, 0, 1, 0, 0};
  for (int i = 0; i < 8; ++i) L.tgt.sp[i] = sp[i];
  L.prop_std = 1.0; L.n_chains = C; L.seed = 42; L.state = state; L.n_steps = n; L.out = out; L.out_n = n; L.accept_total = acc;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f, sum = 0;
  for (int it = 0; it < 7; ++it) {
    L.step_base = it * n;
    cudaEventRecord(e0);
    cudaError_t e = gm::launch_mh_fast(L, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch failed\n"); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it >= 2) { best = ms < best ? ms : best; sum += ms; }
  }
  unsigned long long h; cudaMemcpy(&h, acc, 8, cudaMemcpyDeviceToHost);
  std::vector<double> tail(4); cudaMemcpy(tail.data(), out + (C * n - 2) * 2, 32, cudaMemcpyDeviceToHost);
  printf("%s: best %.3f ms, mean %.3f ms, %.3f of 6543 GB/s, accept %.4f, last %.4f %.4f\n", argv[0], best, sum / 5, (C * n * 16.0 / (sum / 5 * 1e-3)) / 6543.1e9,
         (double)h / (7.0 * C * n), tail[2], tail[3]);
  return 0;
}
