// tools/microbench_write.cu — what a WRITE-ONLY stream reaches on this GPU (the K2 sample write-out is one):
//   (a) cudaMemsetAsync, (b) a grid-stride st.global.cs.v4 kernel, (c) the K2 flush pattern (each warp
//   instruction writes 2 x 256 contiguous bytes 16,000 bytes apart), all over the 16.78 GB of config 2.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void fill_linear(double2* out, size_t n) {
  const double2 v = make_double2(1.0, 2.0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(out + i, v);
}
// chain-major: chain c owns n_steps consecutive 16-byte units; a warp owns 32 chains and writes them 16 steps at a time
__global__ void fill_k2(double2* out, size_t n_chains, int n_steps) {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31, half = lane >> 4, t = lane & 15;
  const double2 v = make_double2(1.0, 2.0);
  double2* base = out + (warp * 32 + half) * (size_t)n_steps + t;
  for (int s0 = 0; s0 + 16 <= n_steps; s0 += 16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) __stcs(base + (size_t)(2 * i) * n_steps + s0, v);
  }
  if (n_steps % 16 && t < n_steps % 16)
    for (int i = 0; i < 16; ++i) __stcs(base + (size_t)(2 * i) * n_steps + (n_steps / 16) * 16, v);
}
// generalised: SEG steps (SEG*16 contiguous bytes) per chain per flush; one warp store instruction covers 512 bytes =
// 512/(SEG*16) chains (SEG <= 32) or part of one chain's segment (SEG > 32).  CS = streaming (evict-first) stores.
template <int SEG, bool CS>
__global__ void fill_seg(double2* out, size_t n_chains, int n_steps) {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const double2 v = make_double2(1.0, 2.0);
  constexpr int LPC = SEG < 32 ? SEG : 32;          // lanes per chain in one instruction
  constexpr int CPI = 32 / LPC;                     // chains per instruction
  constexpr int IPS = SEG / LPC;                    // instructions per chain segment
  const int sub = lane / LPC, t = lane % LPC;
  for (int s0 = 0; s0 + SEG <= n_steps; s0 += SEG) {
#pragma unroll 4
    for (int c = 0; c < 32; c += CPI) {
      double2* p = out + (warp * 32 + c + sub) * (size_t)n_steps + s0 + t;
#pragma unroll
      for (int i = 0; i < IPS; ++i) { if (CS) __stcs(p + i * 32, v); else p[i * 32] = v; }
    }
  }
}
int main() {
  const size_t C = 1048576; const int n = 1000; const size_t units = C * n, bytes = units * 16;
  double2* out; cudaMalloc(&out, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](const char* name, auto fn) {
    float best = 1e9f;
    for (int it = 0; it < 5; ++it) { cudaEventRecord(e0); fn(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms; }
    printf("%-28s %.3f ms  %.0f GB/s written\n", name, best, bytes / (best * 1e-3) / 1e9);
  };
  time("cudaMemsetAsync", [&] { cudaMemsetAsync(out, 0, bytes); });
  time("st.cs.v4 grid-stride 148x8", [&] { fill_linear<<<148 * 8, 256>>>(out, units); });
  time("st.cs.v4 grid-stride 148x16", [&] { fill_linear<<<148 * 16, 512>>>(out, units); });
  time("K2 flush pattern", [&] { fill_k2<<<(unsigned)(C / 128), 128>>>(out, C, n); });
  time("seg 256 B .cs", [&] { fill_seg<16, true><<<(unsigned)(C / 128), 128>>>(out, C, n); });
  time("seg 256 B .wb", [&] { fill_seg<16, false><<<(unsigned)(C / 128), 128>>>(out, C, n); });
  time("seg 512 B .cs", [&] { fill_seg<32, true><<<(unsigned)(C / 128), 128>>>(out, C, 992); });
  time("seg 512 B .wb", [&] { fill_seg<32, false><<<(unsigned)(C / 128), 128>>>(out, C, 992); });
  time("seg 1 KB .cs", [&] { fill_seg<64, true><<<(unsigned)(C / 128), 128>>>(out, C, 960); });
  time("seg 2 KB .cs", [&] { fill_seg<128, true><<<(unsigned)(C / 128), 128>>>(out, C, 896); });
  time("seg 128 B .cs", [&] { fill_seg<8, true><<<(unsigned)(C / 128), 128>>>(out, C, n); });
  printf("(seg 512 B writes 0.992, 1 KB 0.96, 2 KB 0.896 of the bytes: scale their GB/s accordingly)\n");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
