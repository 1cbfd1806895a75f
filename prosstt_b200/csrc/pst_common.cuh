// Shared device helpers: Philox4x32-10, uniform/normal conversions, launch plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/prosstt_b200.h"

namespace pst {

// ---------------------------------------------------------------------------
// host-side plumbing
// ---------------------------------------------------------------------------
extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

int fail_arg(const char *fn, const char *what);
int check_launch(const char *fn);   // cudaGetLastError -> status, counts the launch

#define PST_REQUIRE(cond, fn, what) do { if (!(cond)) return pst::fail_arg(fn, what); } while (0)

constexpr int kNumSM = 148;          // B200: 2 dies x 74 SMs

// stream tags (third Philox counter word): one independent stream per use
enum : uint32_t {
  TAG_WALK_U0 = 0x100, TAG_WALK_V0 = 0x101, TAG_WALK_ETA = 0x102, TAG_WALK_EPS = 0x103,
  TAG_COUNT = 0x200,   // per-(cell,gene) stream of the gamma-Poisson path, block# in word 4
  TAG_QUAD = 0x201,    // one block per (cell, gene quad): the four inversion uniforms
};

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  The key schedule depends only on the seed,
// so the ten round keys are formed once per thread and the rounds are
// 2 x IMAD.WIDE + 2 x LOP3 each.
// ---------------------------------------------------------------------------
struct PhiloxKey {
  uint32_t k0[10], k1[10];
  __host__ __device__ explicit PhiloxKey(uint64_t seed) {
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; ++i) { k0[i] = a; k1[i] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
  }
};

__device__ __forceinline__ uint4 philox(const PhiloxKey &key, uint32_t c0, uint32_t c1,
                                        uint32_t c2, uint32_t c3) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.k0[i];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.k1[i];
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
  }
  return make_uint4(c0, c1, c2, c3);
}

// Same function with the key schedule formed inline from the two seed words: when they are
// kernel parameters the ten bumps live in the uniform datapath (no per-thread key registers,
// no local memory).
__device__ __forceinline__ uint4 philox_s(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                          uint32_t c2, uint32_t c3) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ (k0 + (uint32_t)i * 0x9E3779B9u);
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ (k1 + (uint32_t)i * 0xBB67AE85u);
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
  }
  return make_uint4(c0, c1, c2, c3);
}

// 32 random bits -> float strictly inside (0,1): (w+0.5)*2^-32, top clamped to 1-2^-24
__device__ __forceinline__ float u01(uint32_t w) {
  return fminf(fmaf((float)w, 2.3283064365386963e-10f, 1.1641532182693481e-10f), 0.99999994f);
}
// The uniform of the NB cdf inversion, stretched to (0, 1 + 8e-6].  The fp32 pmf of the inversion
// is exact up to a common factor 1 + eps, |eps| <= ~4e-6 (MUFU lg2/ex2/rcp in P(0)), so its cdf tops
// out anywhere in 1 +- 4e-6.  Stretching u by more than that makes the error one-sided: a u above the
// top of the computed cdf (probability ~8e-6) is detected when the search freezes and the count is
// redrawn, so the accepted draws follow pmf/(1+eps) renormalised, i.e. the exact pmf, and no part of
// the upper tail is cut off.
constexpr float kInversionStretch = 8.0e-6f;
// cdf(0) - u with u = w 2^-32 (1 + stretch) in [0, 1 + 8e-6): no half-step offset (u = 0 is harmless for
// an inversion, X = 0), so the conversion and the subtraction are one I2FP and one FFMA
__device__ __forceinline__ float cdf0_minus_u(float p0, uint32_t w) {
  return fmaf((float)w, -2.3283064365386963e-10f * (1.0f + kInversionStretch), p0);
}
// 53-bit double in [0,1) exactly as numpy's legacy random_sample builds it
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
// standard normal (fp64 Box-Muller, cosine branch) from one Philox block
__device__ __forceinline__ double normal_f64(uint4 r) {
  const double u1 = 1.0 - u53(r.x, r.y);          // (0,1]
  const double u2 = u53(r.z, r.w);
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// fast reciprocal / base-2 log / exp (one MUFU each, no slow-path branches)
__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_fast(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// L2 residency hints: the means table is re-read by every cell (keep), X is written once (stream)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p; asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ float4 ldg_f4_hint(const float *ptr, uint64_t policy) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(policy));
  return v;
}
// 16-byte asynchronous global->shared copy (LDGSTS): a prefetch that holds no registers
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, uint64_t policy) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 16, %2;"
               :: "r"(s), "l"(gmem_src), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void stg_i4_hint(int32_t *ptr, int4 v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;"
               :: "l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy) : "memory");
}

__device__ __forceinline__ void atomic_max_f64(double *addr, double v) {
  unsigned long long *p = reinterpret_cast<unsigned long long *>(addr);
  unsigned long long old = *p;
  while (v > __longlong_as_double((long long)old)) {
    const unsigned long long assumed = old;
    old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

}  // namespace pst
