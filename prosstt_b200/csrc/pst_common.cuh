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

int num_sm();                        // SM count of the current device (148 on a B200), queried once per device
#define PST_SCHED_SLOTS 4096
int sched_slot(void *stream);        // scheduler scratch slot of (current device, stream) for pst_draw_counts

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
// The uniform of the NB cdf inversion, stretched to [0, 1 + 2^-15).  The fp32 pmf of the inversion
// is exact up to a common factor 1 + eps (MUFU lg2/ex2/rcp and the rounding of log2 P(0); |eps| <=
// 1.5e-5 for the parameters routed to the inversion, see nb_inversion_s_max), so its cdf tops out
// anywhere in 1 +- 1.5e-5.  Stretching u by more than that makes the error one-sided: a u above the top
// of the computed cdf is detected (in the tail path, invert_tail) and the count is redrawn,
// so the accepted draws follow pmf (1+eps) renormalised, i.e. the exact pmf, and no part of the
// upper tail is cut off.
constexpr float kInversionStretch = 3.0517578125e-5f;                    // 2^-15
// Philox words whose fp32 image is >= 2^32 (1 - 2^-14) (the top 6.1e-5 of the uniforms) are not
// decided by the fp32 search: their count is inverted in fp64 with 32 more random bits
// (invert_tail).  The route depends on the uniform alone.
constexpr float kTailWord = 4294705152.0f;                               // 2^32 - 2^18
// cdf(0) - u with u = w 2^-32 (1 + stretch): no half-step offset (u = 0 is harmless for an
// inversion, X = 0), so the subtraction is one FFMA on the converted Philox word
__device__ __forceinline__ float cdf0_minus_u(float p0, float w_as_float) {
  return fmaf(w_as_float, -2.3283064365386963e-10f * (1.0f + kInversionStretch), p0);
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
__device__ __forceinline__ float sqrt_fast(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrt_fast(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// natural log / exp / quotient on those: what __logf, __expf and __fdividef compute for normal arguments,
// without their subnormal-argument fix-ups (3-4 extra instructions each); every caller's arguments are
// normal numbers or may be flushed to zero
__device__ __forceinline__ float log_fast(float x) { return lg2_fast(x) * 0.6931471806f; }
__device__ __forceinline__ float exp_fast(float x) { return ex2_fast(x * 1.4426950409f); }
__device__ __forceinline__ float div_fast(float a, float b) { return a * rcp_fast(b); }

// ---------------------------------------------------------------------------
// NB parameterisation of the inversion path, fp32 (count_model.py:156-161 in the gamma-Poisson
// form): theta = alpha mu + beta - 1 (gamma scale), r = mu/theta (gamma shape, scipy's n),
// q = theta/(1+theta) = 1 - p, a = q r, log2 P(0) = -r log2(1+theta).  ONE definition, used by
// the draw kernel's head, its fp64 tail path and the parity hook pst_nb_params_f32.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float nb_shape(float mu, float th) { return mu * rcp_fast(th); }   // r = mu/theta
__device__ __forceinline__ void nb_inversion_params_fast(float mu, float th, float &q, float &a, float &e2) {
  const float t1 = 1.0f + th;
  const float r = nb_shape(mu, th);
  q = th * rcp_fast(t1);
  a = q * r;
  e2 = -r * lg2_fast(t1);                      // log2 P(0) = -r log2(1+theta)
}
// theta -> 0 (Poisson limit): log2 P(0) = -mu log2(e) log1p(theta)/theta by series, exact as theta -> 0
constexpr float kSmallTheta = 0.1f;
__device__ __forceinline__ float nb_log2p0_small_theta(float mu, float th) {
  const float ser = fmaf(th, fmaf(th, fmaf(th, fmaf(th, fmaf(th, fmaf(th, 0.1428571429f, -0.1666666667f), 0.2f),
                                                  -0.25f), 0.3333333333f), -0.5f), 1.0f);
  return -1.4426950409f * mu * ser;
}
__device__ __forceinline__ void nb_inversion_params(float mu, float th, float &q, float &a, float &e2) {
  nb_inversion_params_fast(mu, th, q, a, e2);
  if (th < kSmallTheta) e2 = nb_log2p0_small_theta(mu, th);
}
// Which counts the hybrid sampler inverts (the rest goes to the gamma-Poisson mixture): a function of
// the parameters only, never of the uniforms.
//   mean mu <= 32 and variance mu (1+theta) <= 400   bound the length of a search;
//   shape r <= 48 (or theta < 0.1, the series)        bounds the error of log2 P(0): r |d lg2| <= 48 (2^-22 +
//                                                     2^-24/ln 2) = 1.6e-5, inside the 2^-15 stretch;
//   theta <= 32                                       keeps the tail ratio q below 0.97, so that below the
//                                                     1 - 2^-14 quantile (where the 24-bit fp32 uniform decides)
//                                                     no bin is narrower than ~30 steps of the uniform.
// For a given tree row and gene, mu = M s and theta = alpha M s + (beta-1) grow with the library size s,
// and so do the variance and the shape: the four conditions hold exactly for s <= s_max(M, alpha, beta-1).
// The draw kernel forms s_max once per (chunk, gene) and routes every count with ONE comparison.
// Genes whose parameters are not all positive and finite get s_max = -1: their counts go to the mixture
// queue, whose drain checks the domain (mu > 0, theta > 0) and reports scipy's "Domain error".
constexpr float kInvMuMax = 32.0f, kInvVarMax = 400.0f, kInvShapeMax = 48.0f, kInvThetaMax = 32.0f;
__device__ __forceinline__ float nb_inversion_s_max(float M, float alpha, float bm) {
  if (!(M > 0.f && M < 3.0e38f && alpha >= 0.f && alpha < 3.0e38f && bm > 0.f && bm < 3.0e38f)) return -1.f;
  const float c = alpha * M;                               // theta = c s + bm
  const float b1 = (1.0f + bm) * M;                        // variance = c M s^2 + b1 s
  const float inf = __int_as_float(0x7f800000);
  // (MUFU reciprocals / square root: the thresholds need to be deterministic, not correctly rounded)
  const float rc = rcp_fast(c);                                                    // +inf for c = 0
  float s = kInvMuMax * rcp_fast(M);                                               // mean <= 32
  s = fminf(s, 2.0f * kInvVarMax * rcp_fast(b1 + sqrt_fast(fmaf(b1, b1, 4.0f * kInvVarMax * c * M))));   // variance <= 400
  if (c > 0.f) s = fminf(s, (kInvThetaMax - bm) * rc);                             // theta <= 32
  else if (bm > kInvThetaMax) s = -1.f;
  // shape r = M s / (c s + bm) <= 48  <=>  s (M - 48 c) <= 48 bm;  or theta < 0.1  <=>  s < (0.1 - bm)/c
  const float s_shape = (M > kInvShapeMax * c) ? kInvShapeMax * bm * rcp_fast(M - kInvShapeMax * c) : inf;
  const float s_series = (bm < kSmallTheta) ? (c > 0.f ? (kSmallTheta - bm) * rc : inf) : -1.f;
  return fminf(s, fmaxf(s_shape, s_series));
}
__device__ __forceinline__ bool nb_domain_ok(float mu, float theta) {
  return (mu > 0.f) && (theta > 0.f) && (theta < 3.0e38f) && (mu < 3.0e38f);
}

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
