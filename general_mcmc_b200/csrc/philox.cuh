// philox.cuh — counter-based per-chain random streams (RNG contract: include/gmcmc.h).
// Philox4x32-10 (Salmon et al., SC'11).  Replaces rand::SmallRng / rand_distr ziggurat of the
// reference (generic_hmc.rs:75,92,197; metropolis_hastings.rs:192,313; euclidean.rs:107-113),
// whose bit streams are third-party and unpinned by any reference test.
#pragma once
#include <cstdint>

namespace gm {

struct PhiloxKey { uint32_t k0, k1; };

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, PhiloxKey k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
#else
    uint64_t p0 = (uint64_t)0xD2511F53u * c.x, p1 = (uint64_t)0xCD9E8D57u * c.z;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    c = make_uint4(hi1 ^ c.y ^ k.k0, lo1, hi0 ^ c.w ^ k.k1, lo0);
    k.k0 += 0x9E3779B9u;
    k.k1 += 0xBB67AE85u;
  }
  return c;
}

// Round keys of a launch-constant key, computed once (the key schedule is 20 additions per block otherwise).
struct PhiloxRoundKeys { uint32_t k0[10], k1[10]; };
__host__ __device__ __forceinline__ PhiloxRoundKeys philox_round_keys(PhiloxKey k) {
  PhiloxRoundKeys rk;
#pragma unroll
  for (int r = 0; r < 10; ++r) { rk.k0[r] = k.k0 + (uint32_t)r * 0x9E3779B9u; rk.k1[r] = k.k1 + (uint32_t)r * 0xBB67AE85u; }
  return rk;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxRoundKeys& rk) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ rk.k0[r], lo1, hi0 ^ c.w ^ rk.k1[r], lo0);
  }
  return c;
}

__host__ __device__ __forceinline__ uint4 philox_ctr(uint64_t gchain, uint32_t step, uint32_t stream, uint32_t block) {
  return make_uint4((uint32_t)gchain, (uint32_t)(gchain >> 32), step, (stream << 24) | block);
}

// uniforms in (0,1]
__device__ __forceinline__ float u01(uint32_t r) { return ((float)(r >> 8) + 1.0f) * (1.0f / 16777216.0f); }
__device__ __forceinline__ double u01d(uint32_t hi, uint32_t lo) {
  unsigned long long v = (((unsigned long long)hi << 32) | lo) >> 11;
  return ((double)v + 1.0) * (1.0 / 9007199254740992.0);
}

// Box–Muller.  f32: 4 normals per block; f64: 2 normals per block.
// ACCURATE = false (fast math mode, f32): MUFU approximations (lg2/sqrt/sin/cos.approx, abs. error of the
// normals ~1e-6, far below Monte-Carlo resolution); ACCURATE = true: libdevice logf / sqrtf / sincospif.
template <bool ACCURATE>
__device__ __forceinline__ void normals_from_block(uint4 r, float (&z)[4]) {
  float u0 = u01(r.x), u1 = u01(r.y), u2 = u01(r.z), u3 = u01(r.w);
  if constexpr (ACCURATE) {
    float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
    float s0, c0, s1, c1;
    sincospif(2.0f * u1, &s0, &c0);
    sincospif(2.0f * u3, &s1, &c1);
    z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
  } else {
    // -2 ln u = (-2 ln 2) lg2 u ; angle 2 pi (u - 1/2) in [-pi, pi] where sin/cos.approx are most accurate
    float r0, r1, s0, c0, s1, c1;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(u0));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(u2));
    r0 *= -1.3862943611198906f; r1 *= -1.3862943611198906f;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(r0));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(r1));
    const float a0 = fmaf(u1, 6.283185307179586f, -3.141592653589793f);
    const float a1 = fmaf(u3, 6.283185307179586f, -3.141592653589793f);
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s0) : "f"(a0));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c0) : "f"(a0));
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s1) : "f"(a1));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c1) : "f"(a1));
    z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
  }
}
template <bool ACCURATE>
__device__ __forceinline__ void normals_from_block(uint4 r, double (&z)[2]) {
  double u0 = u01d(r.x, r.y), u1 = u01d(r.z, r.w);
  double rr = sqrt(-2.0 * log(u0));
  double s, c;
  sincospi(2.0 * u1, &s, &c);
  z[0] = rr * c; z[1] = rr * s;
}

// two normals from two 32-bit words (fast path of the 2-D MH kernel; same arithmetic as the first half
// of normals_from_block<false>)
__device__ __forceinline__ void normals_pair_fast(uint32_t w0, uint32_t w1, float& z0, float& z1) {
  float r0, s0, c0;
  const float u0 = u01(w0), u1 = u01(w1);
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(u0));
  r0 *= -1.3862943611198906f;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(r0));
  const float a0 = fmaf(u1, 6.283185307179586f, -3.141592653589793f);
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s0) : "f"(a0));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c0) : "f"(a0));
  z0 = r0 * c0; z1 = r0 * s0;
}

template <class T> struct NormalsPerBlock;
template <> struct NormalsPerBlock<float> { static constexpr int value = 4; };
template <> struct NormalsPerBlock<double> { static constexpr int value = 2; };

template <class T> __device__ __forceinline__ T accept_uniform(uint4 r);
template <> __device__ __forceinline__ float accept_uniform<float>(uint4 r) { return u01(r.x); }
template <> __device__ __forceinline__ double accept_uniform<double>(uint4 r) { return u01d(r.x, r.y); }

}  // namespace gm
