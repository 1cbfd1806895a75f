// The hot loop: per cell x gene negative-binomial draw (prosstt/simulation.py:602-651,
// prosstt/count_model.py:131-161, and numpy's legacy RandomState.negative_binomial behind
// scipy.stats.nbinom.rvs at simulation.py:647-648).
//
//   mu    = means[row_of_cell[n]][g] * scaling[n]
//   theta = alpha[g]*mu + (beta[g]-1)          gamma scale   (scipy p = 1/(1+theta))
//   r     = mu / theta                         gamma shape   (scipy n)
//   X     ~ Poisson( theta * Gamma(r) )
//
// Work item = (cell, gene quad): one thread draws 4 neighbouring genes of one cell and
// writes them with one 128-bit store; a warp writes 512 contiguous bytes of X.  Items are
// a flat range so the grid is load-balanced for any G.  Every draw uses its own Philox
// stream with counter (gene, cell_lo, TAG|cell_hi, block#), so counts do not depend on
// the launch shape, the cell partition or the GPU count.
#include "pst_common.cuh"

namespace pst {

// ---------------------------------------------------------------------------
// per-(cell,gene) word stream
// ---------------------------------------------------------------------------
struct GeneStream {
  const PhiloxKey &key;
  uint32_t c0, c1, c2, blk;
  uint4 buf;
  int left;
  __device__ GeneStream(const PhiloxKey &k, uint32_t gene, int64_t cell)
      : key(k), c0(gene), c1((uint32_t)cell),
        c2((TAG_COUNT << 16) | (uint32_t)((uint64_t)cell >> 32)), blk(0), left(0) {}
  __device__ __forceinline__ uint32_t next() {
    if (left == 0) { buf = philox(key, c0, c1, c2, blk++); left = 4; }
    const uint32_t w = buf.x;
    buf.x = buf.y; buf.y = buf.z; buf.z = buf.w;
    --left;
    return w;
  }
  __device__ __forceinline__ float uniform() { return u01(next()); }
};

// log(k!) : exact table below 16, Stirling series above (abs error < 1e-7 relative to 1)
__device__ __forceinline__ float log_factorial_small(int k) {
  const float tab[16] = {0.f, 0.f, 0.6931471806f, 1.7917594692f, 3.1780538303f, 4.7874917428f,
                         6.5792512120f, 8.5251613611f, 10.6046029027f, 12.8018274801f,
                         15.1044125730f, 17.5023078459f, 19.9872144957f, 22.5521638531f,
                         25.1912211827f, 27.8992713838f};
  return tab[k];
}

// Poisson(lam), lam >= 10: PTRS transformed rejection (Hoermann 1993), the algorithm numpy's
// legacy generator uses for lam >= 10.  The acceptance bound -lam + k log(lam) - log(k!) is
// evaluated as k(log1p(y)-y) - log(sqrt(2 pi k)) - 1/(12k)+..., y = (lam-k)/k, which stays
// accurate in fp32 for lam up to 2^24.
template <class Stream>
__device__ __forceinline__ float poisson_ptrs(float lam, Stream &rng) {
  const float slam = sqrtf(lam);
  const float loglam = __logf(lam);
  const float b = 0.931f + 2.53f * slam;
  const float a = -0.059f + 0.02483f * b;
  const float inv_alpha = 1.1239f + 1.1328f / (b - 3.4f);
  const float vr = 0.9277f - 3.6224f / (b - 2.0f);
  float k = 0.f;
  for (int it = 0; it < 64; ++it) {
    const float U = rng.uniform() - 0.5f;
    const float V = rng.uniform();
    const float us = 0.5f - fabsf(U);
    k = floorf((2.0f * a / us + b) * U + lam + 0.43f);
    if (us >= 0.07f && V <= vr) break;
    if (k < 0.f || (us < 0.013f && V > us)) continue;
    float bound;
    if (k < 16.f) {
      bound = -lam + k * loglam - log_factorial_small((int)k);
    } else {
      const float y = (lam - k) / k;
      const float ik = 1.0f / k;
      bound = k * (log1pf(y) - y) - 0.5f * __logf(6.2831853072f * k) -
              ik * (0.0833333333f - 0.0027777778f * ik * ik);
    }
    if (__logf(V) + __logf(inv_alpha) - __logf(a / (us * us) + b) <= bound) break;
  }
  return k;
}

// Poisson(lam), lam < 10: sequential inversion with one uniform
template <class Stream>
__device__ __forceinline__ float poisson_small(float lam, Stream &rng) {
  const float u = rng.uniform();
  float p = __expf(-lam), cdf = p, k = 0.f;
  while (u > cdf && k < 96.f) {
    k += 1.0f;
    p *= __fdividef(lam, k);
    cdf += p;
  }
  return k;
}

// standard gamma(shape a >= 2/3.. any a>=1 in use) / by Marsaglia-Tsang; returns d*v
template <class Stream>
__device__ __forceinline__ float gamma_mt(float a, Stream &rng) {
  const float d = a - 0.3333333333f;
  const float c = rsqrtf(9.0f * d);
  float v = 1.0f;
  for (int it = 0; it < 64; ++it) {
    // Box-Muller (cosine branch)
    const float u1 = rng.uniform(), u2 = rng.uniform();
    const float x = sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831853072f * u2 - 3.1415926536f);
    const float e = c * x;
    const float t = 1.0f + e;
    if (t <= 0.f) continue;
    v = t * t * t;
    const float u = rng.uniform();
    const float x2 = x * x;
    if (u < 1.0f - 0.0331f * x2 * x2) break;
    // log u < x^2/2 + d(1 - v + log v); for small |e| the right side is the series
    // d e^4 (-3/4 + 3/5 e - 1/2 e^2 + 3/7 e^3 - 3/8 e^4 ...) (cancellation-free)
    float h;
    if (fabsf(e) < 0.1f) {
      const float e2 = e * e;
      h = d * e2 * e2 * (-0.75f + e * (0.6f + e * (-0.5f + e * (0.4285714286f + e * (-0.375f + e * 0.3333333333f)))));
    } else {
      h = 0.5f * x2 + d * (1.0f - v + __logf(v));
    }
    if (__logf(u) < h) break;
  }
  return d * v;
}

// one NB count by the gamma-Poisson mixture
template <class Stream>
__device__ __forceinline__ int nb_gamma_poisson(float mu, float alpha, float bm1, Stream &rng,
                                                uint32_t &flag) {
  const float theta = fmaf(alpha, mu, bm1);
  if (!(mu > 0.f) || !(theta > 0.f) || !(mu < 3.0e38f) || !(theta < 3.0e38f)) {
    flag |= PST_FLAG_DOMAIN;
    return 0;
  }
  const float r = mu / theta;
  float g;
  if (r >= 1.0f) {
    g = gamma_mt(r, rng);
  } else {
    // Gamma(r) = Gamma(r+1) * U^(1/r)
    const float u = rng.uniform();
    g = gamma_mt(r + 1.0f, rng) * __expf(__fdividef(__logf(u), r));
  }
  const float lam = theta * g;
  float k;
  if (lam < 10.f) k = poisson_small(lam, rng);
  else if (lam < 1.6e7f) k = poisson_ptrs(lam, rng);
  else {
    // beyond 2^24 a float cannot hold every integer: normal limit (TV error < 1e-4)
    const float u1 = rng.uniform(), u2 = rng.uniform();
    k = rintf(lam + sqrtf(lam) * sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831853072f * u2 - 3.1415926536f));
  }
  if (k > 2147483520.f) { flag |= PST_FLAG_CLAMPED; return 2147483647; }
  return (int)k;
}

// ---------------------------------------------------------------------------
// kernel v0: straight gamma-Poisson per count
// ---------------------------------------------------------------------------
constexpr int DC_THREADS = 256;

template <bool VEC>
__global__ void __launch_bounds__(DC_THREADS)
draw_counts_gp_kernel(PhiloxKey key, const float *__restrict__ means, int64_t P, int64_t G, int64_t Q,
                      const int32_t *__restrict__ row_of_cell, const float *__restrict__ scaling,
                      const float *__restrict__ alpha, const float *__restrict__ beta_m1,
                      int64_t cell0, int64_t n, int32_t *__restrict__ X, int64_t ldx,
                      uint32_t *__restrict__ flags) {
  const int64_t items = n * Q;
  const int64_t stride = (int64_t)gridDim.x * DC_THREADS;
  uint32_t flag = 0;
  for (int64_t it = (int64_t)blockIdx.x * DC_THREADS + threadIdx.x; it < items; it += stride) {
    const int64_t cell = it / Q;
    const int64_t g0 = (it - cell * Q) * 4;
    const int32_t row = row_of_cell[cell];
    int out[4] = {0, 0, 0, 0};
    if (row < 0 || row >= P) {
      flag |= PST_FLAG_ROW;
    } else {
      const float s = scaling[cell];
      float m[4], a[4], b[4];
      if (VEC) {
        const float4 mv = *reinterpret_cast<const float4 *>(means + (int64_t)row * G + g0);
        const float4 av = *reinterpret_cast<const float4 *>(alpha + g0);
        const float4 bv = *reinterpret_cast<const float4 *>(beta_m1 + g0);
        m[0] = mv.x; m[1] = mv.y; m[2] = mv.z; m[3] = mv.w;
        a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
        b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = g0 + j < G;
          m[j] = ok ? means[(int64_t)row * G + g0 + j] : 1.f;
          a[j] = ok ? alpha[g0 + j] : 0.f;
          b[j] = ok ? beta_m1[g0 + j] : 1.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (VEC || g0 + j < G) {
          GeneStream rng(key, (uint32_t)(g0 + j), cell0 + cell);
          out[j] = nb_gamma_poisson(m[j] * s, a[j], b[j], rng, flag);
        }
      }
    }
    if (VEC) {
      __stcs(reinterpret_cast<int4 *>(X + cell * ldx + g0), make_int4(out[0], out[1], out[2], out[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (g0 + j < G) X[cell * ldx + g0 + j] = out[j];
    }
  }
  if (flag) atomicOr(flags, flag);
}

}  // namespace pst

using namespace pst;

extern "C" int pst_draw_counts(const float *means, int64_t P, int64_t G, const int32_t *row_of_cell,
                               const float *scaling, const float *alpha, const float *beta_m1,
                               uint64_t seed, int64_t cell0, int64_t n, int32_t *X, int64_t ldx,
                               uint32_t *flags, int32_t sampler, void *stream) {
  const char *fn = "pst_draw_counts";
  PST_REQUIRE(P >= 0 && G >= 0 && n >= 0 && cell0 >= 0, fn, "negative size");
  PST_REQUIRE(ldx >= G, fn, "ldx < G");
  PST_REQUIRE(G < (int64_t)1 << 32, fn, "G must be below 2^32");
  PST_REQUIRE(cell0 + n < (int64_t)1 << 48, fn, "cell index must be below 2^48");
  PST_REQUIRE(sampler == PST_SAMPLER_GAMMA_POISSON || sampler == PST_SAMPLER_HYBRID, fn, "unknown sampler");
  if (n == 0 || G == 0) return 0;
  PST_REQUIRE(means && row_of_cell && scaling && alpha && beta_m1 && X && flags, fn, "null pointer");
  const int64_t Q = (G + 3) / 4;
  const bool vec = (G % 4 == 0) && (ldx % 4 == 0) && ((uintptr_t)means % 16 == 0) &&
                   ((uintptr_t)alpha % 16 == 0) && ((uintptr_t)beta_m1 % 16 == 0) &&
                   ((uintptr_t)X % 16 == 0);
  const int64_t items = n * Q;
  int64_t blocks = (items + DC_THREADS - 1) / DC_THREADS;
  const int64_t cap = (int64_t)kNumSM * 8 * 4;         // 4 waves of 8 CTAs per SM, then grid-stride
  if (blocks > cap) blocks = cap;
  const PhiloxKey key(seed);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec)
    draw_counts_gp_kernel<true><<<(unsigned)blocks, DC_THREADS, 0, st>>>(
        key, means, P, G, Q, row_of_cell, scaling, alpha, beta_m1, cell0, n, X, ldx, flags);
  else
    draw_counts_gp_kernel<false><<<(unsigned)blocks, DC_THREADS, 0, st>>>(
        key, means, P, G, Q, row_of_cell, scaling, alpha, beta_m1, cell0, n, X, ldx, flags);
  return check_launch(fn);
}
