// The hot loop: per cell x gene negative-binomial draw (prosstt/simulation.py:602-651,
// prosstt/count_model.py:131-161, and numpy's legacy RandomState.negative_binomial behind
// scipy.stats.nbinom.rvs at simulation.py:647-648).
//
//   mu    = means[row_of_cell[n]][g] * scaling[n]
//   theta = (alpha[g]*means[row][g]) * scaling[n] + (beta[g]-1)     gamma scale   (scipy p = 1/(1+theta))
//   r     = mu / theta                                               gamma shape   (scipy n)
//   X     ~ NB(r, 1/(1+theta))  ==  Poisson( theta * Gamma(r) )
//
// Work item = (cell, gene quad): one thread draws 4 neighbouring genes of one cell and
// writes them with one 128-bit store; a warp writes 512 contiguous bytes of X.  The cells are
// visited grouped by tree row; work is cut into chunks (32 quads x at most 64 cells of ONE row)
// that persistent warps claim from an atomic counter, so the grid is load-balanced for any G and
// any mix of means.  Every uniform is a pure function of (seed, cell, gene or gene quad, draw
// index) through Philox4x32-10, so counts do not depend on launch shape, visiting order, cell
// partition or GPU count.
//
// Two samplers (both exact up to the relative rounding of fp32 pmf terms, see DESIGN.md "sampler
// accuracy"), two kernels with the same work decomposition:
//   PST_SAMPLER_GAMMA_POISSON  draw_counts_mixture_kernel: every count by the mixture the way NumPy's
//                              legacy generator draws it - Marsaglia-Tsang gamma, then Poisson by PTRS
//                              (lam >= 10) or inversion (lam < 10) - as a pipeline of per-warp queues,
//                              one per stage, so that each stage runs 32 entries wide
//   PST_SAMPLER_HYBRID         draw_counts_kernel: direct inversion of the NB cdf with ONE uniform for
//                              small means (branch-free unrolled head + compacted tails; the top 2^-14
//                              of the uniforms finished by tail_fix_kernel with 64 bits against an fp64
//                              cdf), the mixture for large means
// In both, work that would make a warp diverge is pushed to per-warp shared-memory queues and
// executed 32 entries at a time.  Optional (STATS): per-gene sum / sum of squares / zeros of the
// counts (hybrid: accumulated in registers from what the head stores and corrected by the drains).
#include <stdlib.h>
#include <algorithm>
#include "pst_common.cuh"

namespace pst {

__constant__ float c_logfact[16] = {0.f, 0.f, 0.6931471806f, 1.7917594692f, 3.1780538303f,
                                    4.7874917428f, 6.5792512120f, 8.5251613611f, 10.6046029027f,
                                    12.8018274801f, 15.1044125730f, 17.5023078459f, 19.9872144957f,
                                    22.5521638531f, 25.1912211827f, 27.8992713838f};

// ---------------------------------------------------------------------------
// Mixture draw as a restartable state machine (hybrid kernel).  One call = one Marsaglia-Tsang
// attempt (if lambda is not known yet) + up to two PTRS trials, all from two Philox blocks; a
// rejected entry goes back to the queue and is retried with the next block, so a warp never
// spins in a rejection loop while 31 lanes wait.
//   blocks of the per-(cell,gene) stream: 2a   -> gamma attempt a (normal, accept, boost)
//                                          2a+1 -> Poisson attempt a (two PTRS trials / inversion)
// ---------------------------------------------------------------------------
enum : int { MIX_DONE = 0, MIX_RETRY_GAMMA = 1, MIX_RETRY_POISSON = 2 };
struct MixResult { float value; int status; };   // value: count (DONE) or lambda (RETRY_POISSON)

// Poisson(lam), lam >= 10: one trial of PTRS (Hoermann 1993, the transformed-rejection sampler
// NumPy's legacy generator uses for lam >= 10).  The acceptance bound -lam + k log(lam) - log(k!)
// is evaluated as k(log1p(y) - y) - log(sqrt(2 pi k)) - 1/(12k) + 1/(360k^3), y = (lam-k)/k,
// which stays accurate in fp32 up to lam = 2^24.
__device__ __forceinline__ bool ptrs_trial(float lam, float loglam, float a, float b, float log_inv_alpha,
                                           float vr, float U, float V, float &k) {
  const float us = 0.5f - fabsf(U);
  const float ius = rcp_fast(us);
  k = floorf(fmaf(fmaf(2.0f * a, ius, b), U, lam + 0.43f));
  if (us >= 0.07f && V <= vr) return true;
  if (k < 0.f || (us < 0.013f && V > us)) return false;
  float bound;
  if (k < 16.f) {
    bound = -lam + k * loglam - c_logfact[(int)k];
  } else {
    const float ik = rcp_fast(k);
    const float y = (lam - k) * ik;
    // log1p(y) - y without cancellation: series for |y| < 1/4, log1pf beyond
    float l1;
    if (fabsf(y) < 0.25f) {
      const float y2 = y * y;
      l1 = y2 * (-0.5f + y * (0.3333333333f + y * (-0.25f + y * (0.2f + y * (-0.1666666667f + y * (0.1428571429f +
           y * (-0.125f + y * (0.1111111111f + y * (-0.1f + y * 0.0909090909f)))))))));
    } else {
      l1 = log1pf(y) - y;
    }
    bound = k * l1 - 0.5f * log_fast(6.2831853072f * k) - ik * (0.0833333333f - 0.0027777778f * ik * ik);
  }
  return log_fast(V) + log_inv_alpha - log_fast(fmaf(a * ius, ius, b)) <= bound;
}

// One Marsaglia-Tsang attempt at lambda = theta Gamma(r), r = mu/theta (r < 1: Gamma(r) = Gamma(r+1) U^(1/r)),
// from one Philox block: normal (Box-Muller, cosine branch) from w.x, w.y, acceptance uniform w.z, boost w.w.
// Returns false when the attempt is rejected (never with `force`).
__device__ __forceinline__ bool gamma_attempt(float mu, float theta, uint4 w, bool force, float &lam) {
  const float r = div_fast(mu, theta);
  const bool lt1 = r < 1.0f;
  const float shape = lt1 ? r + 1.0f : r;
  const float d = shape - 0.3333333333f;
  const float c = rsqrt_fast(9.0f * d);
  const float rad2 = -2.0f * log_fast(u01(w.x));                    // Box-Muller, cosine branch
  const float xn = rad2 * rsqrt_fast(fmaxf(rad2, 1e-30f)) * __cosf(6.2831853072f * u01(w.y) - 3.1415926536f);
  const float e = c * xn;
  const float t = 1.0f + e;
  const float v = t * t * t;
  const float u = u01(w.z);
  const float x2 = xn * xn;
  bool ok = t > 0.f;
  if (ok && !(u < 1.0f - 0.0331f * x2 * x2)) {
    float h;
    if (fabsf(e) < 0.1f) {
      const float e2 = e * e;
      h = d * e2 * e2 * (-0.75f + e * (0.6f + e * (-0.5f + e * (0.4285714286f + e * (-0.375f + e * 0.3333333333f)))));
    } else {
      h = 0.5f * x2 + d * (1.0f - v + log_fast(v));
    }
    ok = log_fast(u) < h;
  }
  if (!ok && !force) return false;
  const float boost = lt1 ? exp_fast(div_fast(log_fast(u01(w.w)), r)) : 1.0f;   // Gamma(r) = Gamma(r+1) U^(1/r)
  lam = theta * d * fmaxf(v, 0.f) * boost;
  return true;
}

// The same attempt without branches (same arithmetic on the path each lane would have taken, same result):
// in a full warp the squeeze almost never spares every lane the logarithms, and four of these back to
// back (the quad of a lane in the gamma_poisson kernel) interleave their Philox and MUFU latencies.
__device__ __forceinline__ bool gamma_attempt_dense(float mu, float theta, uint4 w, bool force, float &lam) {
  const float r = div_fast(mu, theta);
  const bool lt1 = r < 1.0f;
  const float shape = lt1 ? r + 1.0f : r;
  const float d = shape - 0.3333333333f;
  const float c = rsqrt_fast(9.0f * d);
  const float rad2 = -2.0f * log_fast(u01(w.x));
  const float xn = rad2 * rsqrt_fast(fmaxf(rad2, 1e-30f)) * __cosf(6.2831853072f * u01(w.y) - 3.1415926536f);
  const float e = c * xn;
  const float t = 1.0f + e;
  const float v = t * t * t;
  const float u = u01(w.z);
  const float x2 = xn * xn;
  const float e2 = e * e;
  const float h_series = d * e2 * e2 * (-0.75f + e * (0.6f + e * (-0.5f + e * (0.4285714286f + e * (-0.375f + e * 0.3333333333f)))));
  const float h_log = 0.5f * x2 + d * (1.0f - v + log_fast(v));
  const float h = fabsf(e) < 0.1f ? h_series : h_log;
  const bool ok = (t > 0.f) && ((u < 1.0f - 0.0331f * x2 * x2) || (log_fast(u) < h));
  const float boost = lt1 ? exp_fast(div_fast(log_fast(u01(w.w)), r)) : 1.0f;
  lam = theta * d * fmaxf(v, 0.f) * boost;
  return ok || force;
}

__device__ __noinline__ MixResult mixture_step(float x, float theta, bool have_lambda, int attempt,
                                               uint32_t key0, uint32_t key1, uint32_t gene, int64_t cell) {
  const uint32_t c1 = (uint32_t)cell, c2 = (TAG_COUNT << 16) | (uint32_t)((uint64_t)cell >> 32);
  const bool force = attempt >= 62;               // never reached in practice (p ~ 0.1^62)
  float lam;
  int pa = attempt;                                // Poisson attempt index
  if (!have_lambda) {                              // x = mu: gamma stage
    const uint4 w = philox_s(key0, key1, gene, c1, c2, 2u * (uint32_t)attempt);
    if (!gamma_attempt(x, theta, w, force, lam)) return MixResult{0.f, MIX_RETRY_GAMMA};
    pa = 0;
  } else {
    lam = x;                                       // gamma already accepted
  }
  const uint4 w = philox_s(key0, key1, gene, c1, c2, 2u * (uint32_t)pa + 1u);
  float k;
  if (lam < 10.f) {                                // inversion, one uniform
    if (w.x >= 0xFFF00000u) {
      // top 2^-12 of the uniforms: fp32 cannot resolve the cdf next to 1 (and its uniform has 24 bits), so
      // the far tail is inverted with 32 more random bits against a cdf accumulated in fp64
      const double u = ((double)w.x + ((double)w.y + 0.5) * 2.3283064365386963e-10) * 2.3283064365386963e-10;
      const double dl = (double)lam;
      double p = exp(-dl), cdf = p, kk = 0.0;
      while (u > cdf && kk < 1024.0) { kk += 1.0; p *= dl / kk; cdf += p; }
      k = (float)kk;
    } else {
      const float u = u01(w.x);
      float p = exp_fast(-lam), cdf = p;
      k = 0.f;
      while (u > cdf && k < 96.f) { k += 1.0f; p *= div_fast(lam, k); cdf += p; }
    }
  } else if (lam < 1.6e7f) {                       // PTRS (Hoermann 1993), two trials per block
    const float slam = lam * rsqrt_fast(lam), loglam = log_fast(lam);
    const float b = fmaf(2.53f, slam, 0.931f);
    const float a = fmaf(0.02483f, b, -0.059f);
    const float log_inv_alpha = log_fast(1.1239f + div_fast(1.1328f, b - 3.4f));
    const float vr = 0.9277f - div_fast(3.6224f, b - 2.0f);
    bool ok = ptrs_trial(lam, loglam, a, b, log_inv_alpha, vr, u01(w.x) - 0.5f, u01(w.y), k);
    if (!ok) ok = ptrs_trial(lam, loglam, a, b, log_inv_alpha, vr, u01(w.z) - 0.5f, u01(w.w), k);
    if (!ok && !force) return MixResult{lam, MIX_RETRY_POISSON};
    k = fmaxf(k, 0.f);
  } else {                                         // beyond 2^24: normal limit (TV error < 1e-4)
    k = rintf(lam + sqrtf(lam) * sqrtf(-2.0f * log_fast(u01(w.x))) * __cosf(6.2831853072f * u01(w.y) - 3.1415926536f));   // never hot
  }
  return MixResult{k, MIX_DONE};
}


// ---------------------------------------------------------------------------
// kernel "hybrid": distribution-identical, divergence-aware, warp-autonomous.
//
// Small means (mean <= 32, variance <= 400, theta <= 32, shape <= 48 or theta < 0.1; decided per
//   count as s <= s_max(M, alpha, beta-1), nb_inversion_s_max): inversion of the NB cdf with one
//   uniform u:  P(0) = (1+theta)^-r,  P(k+1) = P(k) (a + q k)/(k+1),
//   q = theta/(1+theta), a = q r,  X = #{k : u > cdf(k)}.
//   Head: KFIX terms fully unrolled and branch-free for the 4 genes of a thread.  The state is
//   t_k = P(k) k! and d_k = cdf(k) - u, so one term costs FFMA (a+qk), FMUL (t), FFMA
//   (d += t/k!, 1/k! an immediate) and one integer op adding the sign bit of d to the count.
//   The four uniforms of a thread come from ONE Philox block keyed (cell, gene quad).
//   Tail: counts still undecided after KFIX terms (u above the cdf) are pushed to a per-warp
//   shared-memory queue and finished 32 at a time with k warp-uniform: HY_STAGE2 more unrolled
//   terms; what is still open then moves to a queue of long searches, drained 32 at a time by a
//   generic loop with a vote every 4 terms.  The tail runs at full warp width whatever the mix of means.
//   Far tail: counts whose Philox word is in the top 2^-14 are listed and finished by
//   tail_fix_kernel (64-bit uniform, fp64 cdf of the same fp32 terms, redraw above the cdf's top).
// Large means are pushed to a third per-warp queue and drawn 32 at a time by the mixture
//   (Marsaglia-Tsang + PTRS) on their own per-(cell, gene) Philox stream; the draw is a
//   restartable step (mixture_step): rejected entries are re-queued, no warp spins in a loop.
// The head writes the quad with one 128-bit store (a partial count in undecided and mixture
//   slots); queue results overwrite them with 4-byte stores after a __syncwarp (same warp, ordered).
// The inversion / mixture route depends on the parameters only, never on the uniforms, so the draw is
//   unbiased; the far-tail route depends on the uniform alone and both sides invert the same cdf.
// Work decomposition and the dynamic chunk scheduler are described at the kernel.
// ---------------------------------------------------------------------------
constexpr int HY_WARPS = 4;                  // warps per CTA
constexpr int HY_THREADS = HY_WARPS * 32;
constexpr int HY_QCAP = 160;                 // 31 carried + 128 new entries, rounded up
constexpr int HY_KMAX = 2048;                // hard bound on inversion terms
#ifndef HY_STAGE2
#define HY_STAGE2 24                         // unrolled terms at the start of the tail
#endif
constexpr float HY_MU_MAX = kInvMuMax;                       // route: see nb_inversion_s_max
#ifndef HY_KFIX_N
#define HY_KFIX_N 10
#endif
constexpr int HY_KFIX = HY_KFIX_N;                           // unrolled head terms
// mean <= 32 keeps P(0) >= e^-32 (|log2 P(0)| <= 46, t_k = P(k) k! inside the fp32 range)
static_assert(HY_MU_MAX <= 32.0f, "the t_k form and the error bound of log2 P(0) assume mean <= 32");

// 1/(k+1) for the open-ended part of the search drain, k = K_TAIL0 + j: read four at a time from the
// constant bank with a warp-uniform index instead of one MUFU.RCP per term
constexpr int HY_TAIL0 = HY_KFIX - 1 + HY_STAGE2;          // first k of the open-ended loop (33)
struct RcpTail { float v[HY_KMAX]; };
__host__ __device__ constexpr RcpTail make_rcp_tail() {
  RcpTail t{};
  for (int j = 0; j < HY_KMAX; ++j) t.v[j] = 1.0f / (float)(HY_TAIL0 + j + 1);
  return t;
}
__constant__ __align__(16) RcpTail c_rcp_tail = make_rcp_tail();

// 1/k!, k = 0..33 (immediates after unrolling; 1/33! = 1.15e-37 is the last normal fp32 value of the series)
#define PST_INV_FACT_TABLE {1.000000000e+00f, 1.000000000e+00f, 5.000000000e-01f, 1.666666667e-01f,               \
    4.166666667e-02f, 8.333333333e-03f, 1.388888889e-03f, 1.984126984e-04f, 2.480158730e-05f, 2.755731922e-06f,   \
    2.755731922e-07f, 2.505210839e-08f, 2.087675699e-09f, 1.605904384e-10f, 1.147074560e-11f, 7.647163732e-13f,   \
    4.779477332e-14f, 2.811457254e-15f, 1.561920697e-16f, 8.220635247e-18f, 4.110317623e-19f, 1.957294106e-20f,   \
    8.896791392e-22f, 3.868170171e-23f, 1.611737571e-24f, 6.446950284e-26f, 2.479596263e-27f, 9.183689864e-29f,   \
    3.279889237e-30f, 1.130996289e-31f, 3.769987629e-33f, 1.216125042e-34f, 3.800390755e-36f, 1.151633562e-37f}

// runtime-indexed copy of 1/k! for the tail path (the kernels use the immediates)
__constant__ float c_inv_fact[34] = PST_INV_FACT_TABLE;

// one term of the inversion in the head's form: t_{k+1} = t_k (a + q k), t_k = P(k) k!
__device__ __forceinline__ float inv_t_next(float t, float a, float q, int k) {
  return t * ((k == 0) ? a : fmaf(q, (float)k, a));
}

// Far upper tail of the inversion (the top 2^-14 of the uniforms, listed by the draw kernel by the
// Philox word alone and finished by tail_fix_kernel) and the redraws.  fp32 cannot resolve the cdf next
// to 1 (spacing 6e-8) and its uniform has 24 bits, so these counts are inverted with a 64-bit uniform -
// the head's word w refined by 32 more random bits, U = (w + (w2 + 1/2) 2^-32) 2^-32, stretched like
// the fp32 uniform (u = U (1 + 2^-15)) - against a cdf accumulated in fp64.  The pmf terms are the
// SAME fp32 numbers the head and the search drain add up (t_k = P(k) k! through k = 33, then P(k)
// itself; same functions, same tables): the two cdfs differ only by the unbiased rounding of the fp32
// accumulation, so they join where the route switches.  (An independent fp64 pmf does not: after ~200
// terms the fp32 terms have drifted by ~5e-7 of cdf, which showed as one displaced bin at 1e9 draws.)
// The cdf tops out at T = 1 + eps like the head's.  u >= T (probability ~2^-15) has no crossing: the
// count is redrawn with a fresh 64-bit uniform over the whole range, so that accepted draws follow
// p(k) / T, the exact pmf up to the relative rounding of its terms.  Random words: blocks 0xffffffff,
// 0xfffffffe, ... of the count's own (cell, gene) stream, which the mixture (blocks 2a, 2a+1, a <= 62)
// never reaches.
constexpr int HY_KMAX64 = 1 << 16;
__device__ __forceinline__ int invert_tail(float mu, float th, uint32_t key0, uint32_t key1, uint32_t gene,
                                           int64_t cell) {
  float q, a, e2;
  nb_inversion_params(mu, th, q, a, e2);
  const float p0 = ex2_fast(e2);
  const uint32_t c1 = (uint32_t)cell, chi = (uint32_t)((uint64_t)cell >> 32);
  const uint4 wq = philox_s(key0, key1, gene >> 2, c1, (TAG_QUAD << 16) | chi, 0u);   // the head's block
  const uint32_t j4 = gene & 3u;
  const uint32_t w = j4 == 0 ? wq.x : j4 == 1 ? wq.y : j4 == 2 ? wq.z : wq.w;
  const double stretch = 1.0 + (double)kInversionStretch;
  int k = 0;
  for (uint32_t attempt = 0; attempt < 16u; ++attempt) {
    const uint4 wt = philox_s(key0, key1, gene, c1, (TAG_COUNT << 16) | chi, 0xffffffffu - attempt);
    const double hi = attempt == 0 ? (double)w : (double)wt.y;
    const double u = (hi + ((double)wt.x + 0.5) * 2.3283064365386963e-10) * 2.3283064365386963e-10 * stretch;
    float t = p0;
    double cdf = (double)p0;
    bool above = false;
    k = 0;
    while (cdf < u && k < HY_TAIL0) {                       // head + stage 2 of the drain: t_k form
      t = inv_t_next(t, a, q, k);
      cdf += (double)t * (double)c_inv_fact[k + 1];
      ++k;
    }
    if (cdf < u) {                                          // the drain's open-ended loop: P(k) form
      float pp = t * c_inv_fact[HY_TAIL0];
      float ak = fmaf(q, (float)HY_TAIL0, a);
      while (cdf < u) {
        // past the mode with a term that cannot move the cdf any more: u lies above the top T
        if (k >= HY_KMAX64 || (pp < 1e-30f && (float)k > mu)) { above = true; break; }
        const int j = k - HY_TAIL0;
        const float rk = j < HY_KMAX ? c_rcp_tail.v[j] : 1.0f / (float)(k + 1);
        pp *= ak * rk;                                      // P(k+1) = P(k) (a + q k)/(k+1)
        cdf += (double)pp;
        ak += q;
        ++k;
      }
    }
    if (!above) break;
  }
  return k;
}

// work-scheduler words of pst_draw_counts: [slot][0] next chunk, [slot][1] warps that have left; one
// slot per (device, stream), see sched_slot().  Zero between launches (the last warp rearms them).
__device__ unsigned int g_sched[PST_SCHED_SLOTS][2];

constexpr int HY_MCAP = HY_QCAP;             // mixture queue: 31 carried + 128 new
constexpr int HY_LCAP = 40;                  // long searches: a full batch plus the few a drained batch adds
struct HyWarpQueues {
  float4 se[HY_QCAP];       // inversion tail: t = P(k) k! at k = KFIX-1, cdf(k)-u, a, q
  int2 sw[HY_QCAP];         //                 where the count goes: (cell, gene)
  float4 le[HY_LCAP];       // long searches (still open after stage 2): P(k), cdf(k)-u, a + q k, q at k = HY_TAIL0
  int2 lw[HY_LCAP];         //                 where the count goes
  float4 ge[HY_MCAP];       // mixture: mu (or lambda once the gamma is accepted), theta, cell, gene
  int ga[HY_MCAP];          //          attempt counter of the current stage | stage << 16
  int fill[4];              // [0] search entries (appended with shared-memory atomics), [1] long searches, [2] mixture entries
};

#ifndef HY_MIN_CTAS
#define HY_MIN_CTAS 7     // 28 warps/SM: up to 72 registers and 29.6 KB of queues per CTA
#endif
#ifndef HY_ENQ_ATOMIC
#define HY_ENQ_ATOMIC 1   // search-queue slots from a shared-memory counter (1) or from ballots (0)
#endif
#ifndef HY_MIX_ATOMIC
#define HY_MIX_ATOMIC 0   // mixture-queue slots: one atomic per lane (1) or ballots (0); measured equal at the bench
                          // depth, ballots better when every count takes the mixture (gamma_poisson sampler)
#endif
#ifndef HY_META2
#define HY_META2 1        // cell metadata of the whole group loaded at the chunk start (1) or every 32 cells in the loop (0)
#endif
#ifndef HY_CHUNK_CELLS
#define HY_CHUNK_CELLS 64                    // most cells per chunk (x 32 quads = 8192 counts); all of one tree row
#endif

// Work decomposition.  The cells of a launch are visited grouped by tree row (counting sort,
// group_* kernels below): a GROUP is a run of at most HY_CHUNK_CELLS cells that share their row, a
// CHUNK is one group times one strip of 32 gene quads (one quad per lane, 128 genes = 512 contiguous
// bytes of each X row).  Inside a chunk everything that depends on (row, gene) is constant:
// the means quad M, c = alpha M, beta-1 and the largest library size s for which the gene is routed to
// the inversion (nb_inversion_s_max) are formed once per chunk; per cell only its id and its library
// size s are read (by the lanes, 32 cells per request, broadcast with shuffles), mu = M s,
// theta = c s + beta-1, and the route is one comparison s <= s_max.  Chunks are handed out by an
// atomic counter, so warps whose queues drain more often simply take fewer of them; consecutive
// chunk ids are adjacent strips of the same cells (neighbouring warps complete the same X rows
// together), and concurrently running warps share means rows in L2.
#ifndef HY_STATS_CTAS
#define HY_STATS_CTAS 7   // CTAs per SM of the instantiation with fused per-gene summaries
#endif
template <int KFIX, bool VEC, bool STATS>
__global__ void __launch_bounds__(HY_THREADS, STATS ? HY_STATS_CTAS : HY_MIN_CTAS)
draw_counts_kernel(const __grid_constant__ PhiloxKey key, const float *__restrict__ means, uint32_t G, uint32_t Q,
                   const float *__restrict__ scaling, const float *__restrict__ alpha,
                   const float *__restrict__ beta_m1, int64_t cell0, int32_t *__restrict__ X, uint32_t ldx,
                   uint32_t *__restrict__ flags, const uint32_t *__restrict__ hdr,
                   const int32_t *__restrict__ order, const uint4 *__restrict__ groups, int sched,
                   uint32_t *__restrict__ tail, uint32_t tail_cap, unsigned long long *__restrict__ gene_sum,
                   unsigned long long *__restrict__ gene_sumsq, unsigned long long *__restrict__ gene_zeros) {
  __shared__ HyWarpQueues queues[HY_WARPS];
  __shared__ unsigned int stat_lut[STATS ? 16 : 1];            // c -> [c == 0] | c << 7 | c^2 << 17, c <= KFIX
  if constexpr (STATS) {
    static_assert(KFIX <= 15 && KFIX * HY_CHUNK_CELLS < 1024 && KFIX * KFIX * HY_CHUNK_CELLS < 32768 &&
                  HY_CHUNK_CELLS < 128, "packed per-gene accumulators");
    if (threadIdx.x < 16) {
      const unsigned c = threadIdx.x;
      stat_lut[c] = (c == 0u ? 1u : 0u) | (c << 7) | ((c * c) << 17);
    }
    __syncthreads();
  }
  HyWarpQueues &wq = queues[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint32_t n_strips = (Q + 31u) / 32u;
  const uint64_t n_chunks = (uint64_t)hdr[1] * n_strips;      // hdr[1] = number of groups (group_table_kernel)
  const int64_t n_warps = (int64_t)gridDim.x * HY_WARPS;
  const uint32_t key0 = key.k0[0], key1 = key.k1[0];
  constexpr float inv_fact[34] = PST_INV_FACT_TABLE;
  static_assert(KFIX >= 2 && KFIX - 1 + HY_STAGE2 <= 33, "t_k = P(k) k! must stay in fp32 range through stage 2");
  int ns = 0, ng = 0;                         // queue fill, warp-uniform
  uint32_t flag = 0;
  const uint64_t keep = l2_policy_evict_last();
  if (lane < 4) wq.fill[lane] = 0;
  __syncwarp();


  // Write a finished search.  With fused summaries (STATS) the head has summed the partial count KFIX for
  // this element: the difference is added now.
  auto finish_search = [&](int2 w, int cn) {
    X[(uint64_t)(uint32_t)w.x * ldx + (uint32_t)w.y] = cn;
    if constexpr (STATS) {
      atomicAdd(gene_sum + w.y, (unsigned long long)(cn - KFIX));
      atomicAdd(gene_sumsq + w.y, (unsigned long long)cn * (unsigned long long)cn - (unsigned long long)(KFIX * KFIX));
    }
  };
  // the open-ended part of a search, 32 long searches at a time: P(k) form from k = HY_TAIL0, a + q k
  // advanced by one FADD per term, 1/(k+1) from the constant bank, a vote every 4 terms
  auto drain_long = [&](int first, int cnt) {
    const bool act = lane < cnt;
    const float4 st = act ? wq.le[first + lane] : make_float4(0.f, 1.f, 0.f, 0.f);
    const int2 w = act ? wq.lw[first + lane] : make_int2(0, 0);
    float pp = st.x, dd = st.y, ak = st.z;
    const float qq = st.w;
    int cn = KFIX + HY_STAGE2;
    for (int it = 0; it < HY_KMAX / 4; ++it) {
      if (!__any_sync(0xffffffffu, dd < 0.f)) break;          // every cdf has passed its u
      float rk[4];
      if constexpr (KFIX == HY_KFIX) {                        // 1/(k+1) for the next four k: one constant load
        const float4 r4 = reinterpret_cast<const float4 *>(c_rcp_tail.v)[it];
        rk[0] = r4.x; rk[1] = r4.y; rk[2] = r4.z; rk[3] = r4.w;
      } else {                                                // developer builds with another head length
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) rk[s4] = rcp_fast((float)(KFIX + HY_STAGE2 + 4 * it + s4));
      }
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        pp *= ak * rk[s4];                                     // P(k+1) = P(k) (a + q k)/(k+1)
        dd += pp;                                              // cdf(k+1) - u
        cn += (int)(__float_as_uint(dd) >> 31);
        ak += qq;
      }
    }
    // every u handled here lies below 1 - 2^-15 < T, the top of the computed cdf, and cdf - u is
    // accumulated (exact near the crossing), so each search ends; HY_KMAX only bounds the loop
    if (act) finish_search(w, cn);
  };
  // finish up to 32 queued inversions at full warp width; k is warp-uniform
  auto drain_search = [&](int first, int cnt) {
    const bool act = lane < cnt;
    const int e = first + lane;
    const float4 st = act ? wq.se[e] : make_float4(0.f, 1.f, 0.f, 0.f);
    float tt = st.x, dd = st.y;                            // t = P(KFIX-1) (KFIX-1)!
    const float aa = st.z, qq = st.w;
    int cn = KFIX;
    // second stage: HY_STAGE2 further terms in the head's form t_k = P(k) k! with compile-time k
    // (4 instructions per term instead of 8 in the generic loop below)
#pragma unroll
    for (int s2 = 0; s2 < HY_STAGE2; ++s2) {
      const int k = KFIX - 1 + s2;
      tt = inv_t_next(tt, aa, qq, k);                        // t_{k+1}
      dd = fmaf(tt, inv_fact[k + 1], dd);                    // cdf(k+1) - u
      cn += (int)(__float_as_uint(dd) >> 31);
    }
    const int2 w = act ? wq.sw[e] : make_int2(0, 0);
    // Searches still open after stage 2 (about 6 % of the entries, but 86 % of the batches hold one) move on
    // to the queue of long searches, which is drained 32 at a time: the open-ended loop then runs at full
    // width instead of once per batch for one or two lanes.
    const bool open = act && (dd < 0.f);
    const unsigned mo = __ballot_sync(0xffffffffu, open);
    if (act && !open) finish_search(w, cn);
    if (mo) {
      int nl = *(volatile int *)&wq.fill[1];
      if (nl + __popc(mo) > HY_LCAP) { drain_long(0, nl); nl = 0; __syncwarp(); }   // rare: make room
      if (open) {
        const int l = nl + __popc(mo & lt_mask);
        wq.le[l] = make_float4(tt * inv_fact[KFIX - 1 + HY_STAGE2],            // back to P(k)
                               dd, fmaf(qq, (float)(KFIX - 1 + HY_STAGE2), aa), qq);
        wq.lw[l] = w;
      }
      nl += __popc(mo);
      __syncwarp();
      if (nl >= 32) { nl -= 32; drain_long(nl, 32); }
      if (lane == 0) wq.fill[1] = nl;
      __syncwarp();
    }
  };
  // one mixture step for up to 32 queued entries; rejected entries go back to the queue (returns
  // how many), so the warp never spins in a rejection loop
  auto drain_mixture = [&](int first, int cnt) -> int {
    const bool act = lane < cnt;
    float4 g = make_float4(1.f, 1.f, 0.f, 0.f);
    int ga = 0;                                     // attempt | stage << 16 (stage 1: x holds lambda)
    if (act) { g = wq.ge[first + lane]; ga = wq.ga[first + lane]; }
    __syncwarp();                                   // every entry is read before slots are reused
    const int ccell = __float_as_int(g.z), gene = __float_as_int(g.w);
    const bool have_lambda = (ga >> 16) != 0;
    const int att = ga & 0xffff;
    int status = MIX_DONE;
    float value = 0.f;
    if (act) {
      if (!have_lambda && !nb_domain_ok(g.x, g.y)) {
        flag |= PST_FLAG_DOMAIN;                    // count stays 0
      } else {
        const MixResult res = mixture_step(g.x, g.y, have_lambda, att, key0, key1, (uint32_t)gene, cell0 + ccell);
        status = res.status;
        value = res.value;
      }
      if (status == MIX_DONE) {
        int val = (int)value;
        if (value > 2147483520.f) { val = 2147483647; flag |= PST_FLAG_CLAMPED; }
        X[(uint64_t)(uint32_t)ccell * ldx + (uint32_t)gene] = val;
        if constexpr (STATS) {
          if (val > 0) {                                 // the head summed a zero for this element
            atomicAdd(gene_sum + gene, (unsigned long long)val);
            atomicAdd(gene_sumsq + gene, (unsigned long long)val * (unsigned long long)val);
            atomicAdd(gene_zeros + gene, ~0ull);             // minus one
          }
        }
      }
    }
    const bool again = act && status != MIX_DONE;
    const unsigned m = __ballot_sync(0xffffffffu, again);
    if (again) {
      const int e = first + __popc(m & lt_mask);
      // a rejected gamma keeps (mu, theta); an accepted gamma with rejected Poisson trials keeps
      // lambda (stage 1); either way the next attempt of that stage uses the next Philox block
      wq.ge[e] = (status == MIX_RETRY_GAMMA) ? g : make_float4(value, g.y, g.z, g.w);
      wq.ga[e] = (status == MIX_RETRY_GAMMA) ? att + 1 : ((1 << 16) | (have_lambda ? att + 1 : 1));
    }
    __syncwarp();
    return __popc(m);
  };

  for (;;) {
    unsigned claimed = 0;
    if (lane == 0) claimed = atomicAdd(&g_sched[sched][0], 1u);
    claimed = __shfl_sync(0xffffffffu, claimed, 0);
    if ((uint64_t)claimed >= n_chunks) break;                 // no work left
    const uint32_t grp = claimed / n_strips;
    const uint32_t strip = claimed - grp * n_strips;
    const uint4 gd = __ldg(groups + grp);                     // first position in `order`, cells, tree row
    const uint32_t pos0 = gd.x;
    const int n_cells = (int)gd.y;
    // lanes past the last quad of the last strip repeat quad Q-1: same key, same counts, same
    // addresses, so the duplicates are harmless and no per-count "lane is live" predicate is needed
    const uint32_t quad = min(strip * 32u + (uint32_t)lane, Q - 1u);
    const bool lane_first = strip * 32u + (uint32_t)lane < Q;     // not one of the repeated edge lanes
    const uint32_t g0 = quad * 4u;
    unsigned int acc[4] = {0u, 0u, 0u, 0u};             // STATS: packed per-gene accumulators of this chunk
    // per-(row, gene) constants of this lane's quad, once per chunk
    float m[4], c[4], bm[4], smax[4];
    {
      float al[4];
      const float *mrow = means + (uint64_t)gd.z * G + g0;
      if (VEC) {
        const float4 av = __ldg(reinterpret_cast<const float4 *>(alpha + g0));
        const float4 bv = __ldg(reinterpret_cast<const float4 *>(beta_m1 + g0));
        const float4 mv = ldg_f4_hint(mrow, keep);
        al[0] = av.x; al[1] = av.y; al[2] = av.z; al[3] = av.w;
        bm[0] = bv.x; bm[1] = bv.y; bm[2] = bv.z; bm[3] = bv.w;
        m[0] = mv.x; m[1] = mv.y; m[2] = mv.z; m[3] = mv.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = g0 + j < G;
          al[j] = ok ? alpha[g0 + j] : 0.f;
          bm[j] = ok ? beta_m1[g0 + j] : 1.f;
          m[j] = ok ? mrow[j] : 1.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        c[j] = al[j] * m[j];
        smax[j] = nb_inversion_s_max(m[j], al[j], bm[j]);
      }
    }
    // theta = c s + (beta-1) >= beta-1 on the inversion route, so the Poisson-limit series below can only
    // be needed in chunks where some gene has beta-1 < 0.1: one vote per chunk
    const bool chunk_small_theta =
        __any_sync(0xffffffffu, fminf(fminf(bm[0], bm[1]), fminf(bm[2], bm[3])) < kSmallTheta);
    // Per-cell metadata (cell id, library size) is the same for every lane: lane l loads it for cell
    // (first + l) of the group with one coalesced request per 32 cells and the loop broadcasts it with
    // shuffles.  A library size that is not positive and finite is flagged and sampled as NaN (every
    // count of the cell goes to the mixture queue, whose drain leaves 0).
#if HY_META2
    // A group has at most 64 cells: lane l holds the metadata of cells l and 32 + l, both loaded here, so the
    // cell loop contains no load (the every-32-cells reload used to sit in it as 16 predicated-off instructions)
    static_assert(HY_CHUNK_CELLS <= 64, "two metadata registers per lane cover a group");
    int32_t meta_cell, meta_cell_hi; float meta_s, meta_s_hi;
    auto load_meta = [&](int i, int32_t &mc, float &ms) {
      mc = order[pos0 + (uint32_t)(i < n_cells ? i : n_cells - 1)];
      ms = scaling[mc];
      if (!(ms > 0.f && ms < 3.0e38f)) { flag |= PST_FLAG_DOMAIN; ms = __int_as_float(0x7fc00000); }
    };
    load_meta(lane, meta_cell, meta_s);
    load_meta(32 + lane, meta_cell_hi, meta_s_hi);
    for (int ci = 0; ci < n_cells; ++ci) {
      const int src = ci & 31;
      const bool upper = ci >= 32;
      const int32_t cell = __shfl_sync(0xffffffffu, upper ? meta_cell_hi : meta_cell, src);
      const float s = __shfl_sync(0xffffffffu, upper ? meta_s_hi : meta_s, src);
#else
    int32_t meta_cell = 0; float meta_s = 1.f;
    auto load_meta = [&](int first) {                       // cells first .. first+31 of this group
      const int i = first + lane;
      meta_cell = order[pos0 + (uint32_t)(i < n_cells ? i : n_cells - 1)];
      meta_s = scaling[meta_cell];
      if (!(meta_s > 0.f && meta_s < 3.0e38f)) { flag |= PST_FLAG_DOMAIN; meta_s = __int_as_float(0x7fc00000); }
    };
    load_meta(0);
    for (int ci = 0; ci < n_cells; ++ci) {
      const int src = ci & 31;
      const int32_t cell = __shfl_sync(0xffffffffu, meta_cell, src);
      const float s = __shfl_sync(0xffffffffu, meta_s, src);
      if (src == 31) load_meta(ci + 1);                     // warp-uniform branch, every 32 cells
#endif

      // ---- this cell's quad
      float t[4], d[4], a[4], q[4], mu[4], th[4];
      int cnt[4];
      bool small[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mu[j] = m[j] * s;
        th[j] = fmaf(c[j], s, bm[j]);
        small[j] = s <= smax[j];                    // the inversion's route (false on NaN)
      }
      {
        float e2[4];
        const int64_t gcell = cell0 + cell;
        const uint4 rnd = philox(key, quad, (uint32_t)gcell,
                                 (TAG_QUAD << 16) | (uint32_t)((uint64_t)gcell >> 32), 0u);
        float fw[4] = {(float)rnd.x, (float)rnd.y, (float)rnd.z, (float)rnd.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) nb_inversion_params_fast(mu[j], th[j], q[j], a[j], e2[j]);
        // theta -> 0 (Poisson limit): log1p(theta)/theta by series, exact as theta -> 0
        if (chunk_small_theta && fminf(fminf(th[0], th[1]), fminf(th[2], th[3])) < kSmallTheta) {
#pragma unroll
          for (int j = 0; j < 4; ++j) e2[j] = (th[j] < kSmallTheta) ? nb_log2p0_small_theta(mu[j], th[j]) : e2[j];
        }
        // The top 2^-14 of the uniforms is not decided in fp32: those counts (chosen by the Philox word
        // alone) are appended with their parameters to the launch's tail list and finished by
        // tail_fix_kernel; the head sees them as decided (their word is replaced by one that gives
        // cdf(0) - u > 0: count 0 until the fix-up writes it)
        if (fmaxf(fmaxf(fw[0], fw[1]), fmaxf(fw[2], fw[3])) >= kTailWord) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (small[j] && (fw[j] >= kTailWord) && (VEC || g0 + j < G)) {
              if (lane_first) {                          // repeated edge lanes leave the listing to their twin
                const unsigned slot = atomicAdd(tail, 1u);
                if (slot < tail_cap)
                  reinterpret_cast<uint4 *>(tail + 4)[slot] =
                      make_uint4((uint32_t)cell, g0 + j, __float_as_uint(mu[j]), __float_as_uint(th[j]));
                else
                  flag |= PST_FLAG_SCRATCH;
              }
              fw[j] = -4.0e9f;
            }
          }
        }
        // P(0) and cdf(0) - u for the inversion lanes
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          t[j] = ex2_fast(e2[j]);                     // P(0)
          d[j] = cdf0_minus_u(t[j], fw[j]);           // cdf(0) - u, one FFMA
          cnt[j] = (int)(__float_as_uint(d[j]) >> 31);
        }
      }
      // large means: queue them for the mixture now, so mu/theta are dead during the head.  Many
      // (cell, strip) pairs have none: one vote skips the block
      if (__any_sync(0xffffffffu, !(small[0] && small[1] && small[2] && small[3]))) {
#if HY_MIX_ATOMIC
        // a lane reserves the slots of its (up to four) entries with one shared-memory atomic
        bool tm[4];
        int nm = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) { tm[j] = !small[j] && (VEC || g0 + j < G) && (!STATS || lane_first); nm += tm[j] ? 1 : 0; }
        if (nm) {
          int e = atomicAdd(&wq.fill[2], nm);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (tm[j]) {
              wq.ge[e] = make_float4(mu[j], th[j], __int_as_float((int)cell), __int_as_float((int)(g0 + j)));
              wq.ga[e] = 0;
              ++e;
            }
          }
        }
        __syncwarp();
        ng = *(volatile int *)&wq.fill[2];
#else
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool to_mix = !small[j] && (VEC || g0 + j < G) && (!STATS || lane_first);
          const unsigned mg = __ballot_sync(0xffffffffu, to_mix);
          if (to_mix) {
            const int e = ng + __popc(mg & lt_mask);
            wq.ge[e] = make_float4(mu[j], th[j], __int_as_float((int)cell), __int_as_float((int)(g0 + j)));
            wq.ga[e] = 0;
          }
          ng += __popc(mg);
        }
#endif
      }
      // head of the inversion: terms 1..KFIX-1, branch-free, k compile-time
      {
#pragma unroll
        for (int k = 0; k < KFIX - 1; ++k) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            t[j] = inv_t_next(t[j], a[j], q[j], k);                // t_{k+1} = P(k+1) (k+1)!
            d[j] = fmaf(t[j], inv_fact[k + 1], d[j]);              // cdf(k+1) - u
            cnt[j] += (int)(__float_as_uint(d[j]) >> 31);
          }
        }
      }
      // store the quad: undecided and mixture slots hold a partial count until their queue entry is
      // drained, which always writes them (the __syncwarp below orders this store before the drain's)
      {
        int32_t *dst = X + ((uint64_t)(uint32_t)cell * ldx + g0);
        if constexpr (STATS) {
          // Per-gene summaries fused into the draw (the notebooks' X.sum(axis=0), per-gene variance and zero
          // fraction; pst_count_stats as a second pass costs a quarter of the draw).  What the head stores is
          // summed here: final counts 0..KFIX-1, the partial value KFIX for a queued search and 0 for a
          // mixture or far-tail slot; the drains and tail_fix_kernel add the difference when the element
          // becomes final.  c, c^2 and [c == 0] of at most HY_CHUNK_CELLS head values fit one 32-bit word.
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            cnt[j] = small[j] ? cnt[j] : 0;
            acc[j] += stat_lut[cnt[j]];
          }
        }
        if (VEC) {
          *reinterpret_cast<int4 *>(dst) = make_int4(cnt[0], cnt[1], cnt[2], cnt[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (g0 + j < G) dst[j] = cnt[j];
        }
      }
      // enqueue the undecided inversions
      {
#if HY_ENQ_ATOMIC
        // slots come from a shared-memory counter
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if ((VEC || g0 + j < G) && small[j] && (d[j] < 0.f) && (!STATS || lane_first)) {
            const int e = atomicAdd(&wq.fill[0], 1);
            wq.se[e] = make_float4(t[j], d[j], a[j], q[j]);     // t = P(KFIX-1) (KFIX-1)!, rescaled in the drain
            wq.sw[e] = make_int2((int)cell, (int)(g0 + j));
          }
        }
        __syncwarp();
        ns = *(volatile int *)&wq.fill[0];
#else
        // slots from ballots (all lanes take part)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool push = (VEC || g0 + j < G) && small[j] && (d[j] < 0.f) && (!STATS || lane_first);
          const unsigned ms = __ballot_sync(0xffffffffu, push);
          if (push) {
            const int e = ns + __popc(ms & lt_mask);
            wq.se[e] = make_float4(t[j], d[j], a[j], q[j]);     // t = P(KFIX-1) (KFIX-1)!, rescaled in the drain
            wq.sw[e] = make_int2((int)cell, (int)(g0 + j));
          }
          ns += __popc(ms);
        }
        __syncwarp();
#endif
      }
      if (ns >= 32 || ng >= 32) {
        while (ns >= 32) { ns -= 32; drain_search(ns, 32); }
        while (ng >= 32) { ng -= 32; ng += drain_mixture(ng, 32); }
        __syncwarp();                               // every lane has read the counters before lane 0 rewrites them
#if HY_ENQ_ATOMIC
        if (lane == 0) wq.fill[0] = ns;
#endif
#if HY_MIX_ATOMIC
        if (lane == 0) wq.fill[2] = ng;
#endif
        __syncwarp();
      }
    }
    if constexpr (STATS) {
      // flush the chunk's per-gene accumulators: zeros in bits 0-6, sum in bits 7-16, sum of squares above
      if (lane_first) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (VEC || g0 + j < G) {
            atomicAdd(gene_zeros + g0 + j, (unsigned long long)(acc[j] & 0x7fu));
            atomicAdd(gene_sum + g0 + j, (unsigned long long)((acc[j] >> 7) & 0x3ffu));
            atomicAdd(gene_sumsq + g0 + j, (unsigned long long)(acc[j] >> 17));
          }
        }
      }
    }
  }
  if (ns > 0) drain_search(0, ns);
  __syncwarp();
  {
    const int nl = *(volatile int *)&wq.fill[1];
    if (nl > 0) drain_long(0, nl);
  }
  while (ng > 0) { const int cdone = ng < 32 ? ng : 32; ng -= cdone; ng += drain_mixture(ng, cdone); }
  if (flag) atomicOr(flags, flag);
  // the last warp to leave rearms the scheduler words for the next launch
  if (lane == 0) {
    const unsigned done = atomicAdd(&g_sched[sched][1], 1u);
    if (done == (unsigned)n_warps - 1u) { g_sched[sched][0] = 0u; g_sched[sched][1] = 0u; __threadfence(); }
  }
}

// ---------------------------------------------------------------------------
// kernel "gamma_poisson": every count by the mixture NumPy's legacy generator draws it from
// (Marsaglia-Tsang gamma, then Poisson by inversion below lambda = 10 and by PTRS above), organised as
// a pipeline of three per-warp queues so that every stage runs 32 entries wide:
//   gamma stage     attempt 0 of a count runs straight from the registers of the lane that owns it (all
//                   32 lanes, no queue); an accepted lambda is routed by its size to one of the two
//                   Poisson queues, a rejected attempt (~5 %) goes to the queue of gamma retries
//   small lambda    inversion with one uniform, MX_PS_TERMS terms unrolled and branch-free in the form
//                   t_k = lambda^k e^-lambda, d_k = cdf(k) - u (3 instructions per term); the few still
//                   open continue in a loop with a vote every 4 terms; top 2^-12 of the uniforms in fp64
//   large lambda    PTRS, two trials per Philox block; rejected entries return to the queue
// The Philox blocks of a count are the ones mixture_step uses (2a: gamma attempt a, 2a+1: Poisson attempt
// a of the (cell, gene) stream), so the gamma and PTRS stages draw the very numbers the hybrid kernel's
// mixture path would.  Work decomposition, scheduler and row grouping as in draw_counts_kernel.  Every
// count is written exactly once, by the stage that finishes it (4-byte stores; the entries of a batch
// are neighbouring lanes' genes of one cell, 16 bytes apart, and the other three genes of each quad
// follow within the same iteration: the sectors are completed in L2).
// ---------------------------------------------------------------------------
#ifndef MX_PS_TERMS
#define MX_PS_TERMS 12                       // unrolled terms of the small-lambda inversion (8 / 12 / 16 measured)
#endif
#ifndef MX_MIN_CTAS
#define MX_MIN_CTAS 7
#endif
#ifndef MX_PTRS_ONE_TRIAL
#define MX_PTRS_ONE_TRIAL 1                  // one PTRS trial per visit of the queue (0: two, the second one divergent)
#endif
#ifndef MX_DENSE_GAMMA
#define MX_DENSE_GAMMA 1                     // branch-free Marsaglia-Tsang attempt (gamma_attempt_dense)
#endif
constexpr int MX_CAP = 160;                  // 31 carried + 128 new entries per cell iteration, rounded up
enum : int { MX_Q_GAMMA = 0, MX_Q_SMALL = 1, MX_Q_LARGE = 2 };
struct MxWarpQueues {
  float4 q[3][MX_CAP];         // gamma retries: mu, theta, cell, gene | small lambda: lambda, cell, gene, - |
                               // PTRS: lambda, cell, gene, attempt
  unsigned char g_att[MX_CAP]; // attempt number of a gamma retry (1..62)
};

template <bool VEC, bool STATS>
__global__ void __launch_bounds__(HY_THREADS, MX_MIN_CTAS)
draw_counts_mixture_kernel(uint32_t key0, uint32_t key1, const float *__restrict__ means, uint32_t G, uint32_t Q,
                           const float *__restrict__ scaling, const float *__restrict__ alpha,
                           const float *__restrict__ beta_m1, int64_t cell0, int32_t *__restrict__ X, uint32_t ldx,
                           uint32_t *__restrict__ flags, const uint32_t *__restrict__ hdr,
                           const int32_t *__restrict__ order, const uint4 *__restrict__ groups, int sched,
                           unsigned long long *__restrict__ gene_sum, unsigned long long *__restrict__ gene_sumsq,
                           unsigned long long *__restrict__ gene_zeros) {
  __shared__ MxWarpQueues queues[HY_WARPS];
  MxWarpQueues &wq = queues[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint32_t n_strips = (Q + 31u) / 32u;
  const uint64_t n_chunks = (uint64_t)hdr[1] * n_strips;
  const int64_t n_warps = (int64_t)gridDim.x * HY_WARPS;
  constexpr float inv_fact[34] = PST_INV_FACT_TABLE;
  static_assert(MX_PS_TERMS >= 4 && MX_PS_TERMS <= 24, "lambda^k e^-lambda, lambda < 10, must stay in fp32 range");
  int ng = 0, ns = 0, nl = 0;                // queue fills (gamma retries, small lambda, PTRS), warp-uniform
  uint32_t flag = 0;
  const uint64_t keep = l2_policy_evict_last();

  // a finished count: the only store this element ever gets
  auto write_count = [&](int cell, int gene, float value) {
    int val = (int)value;
    if (value > 2147483520.f) { val = 2147483647; flag |= PST_FLAG_CLAMPED; }
    X[(uint64_t)(uint32_t)cell * ldx + (uint32_t)gene] = val;
    if constexpr (STATS) {
      if (val > 0) {
        atomicAdd(gene_sum + gene, (unsigned long long)val);
        atomicAdd(gene_sumsq + gene, (unsigned long long)val * (unsigned long long)val);
      } else {
        atomicAdd(gene_zeros + gene, 1ull);
      }
    }
  };
  // queue appends: every lane calls, `p` says whether it has an entry
  auto push_large = [&](bool p, float lam, int cell, int gene, int att) {
    const unsigned m = __ballot_sync(0xffffffffu, p);
    if (p) wq.q[MX_Q_LARGE][nl + __popc(m & lt_mask)] =
        make_float4(lam, __int_as_float(cell), __int_as_float(gene), __int_as_float(att));
    nl += __popc(m);
  };
  // one Marsaglia-Tsang attempt (Philox block 2 att of the count's own stream) ...
  auto gamma_draw = [&](float mu, float th, int cell, int gene, int att, float &lam) -> bool {
    const int64_t gcell = cell0 + cell;
    const uint4 w = philox_s(key0, key1, (uint32_t)gene, (uint32_t)gcell,
                             (TAG_COUNT << 16) | (uint32_t)((uint64_t)gcell >> 32), 2u * (uint32_t)att);
#if MX_DENSE_GAMMA
    return gamma_attempt_dense(mu, th, w, att >= 62, lam);
#else
    lam = 0.f;
    return gamma_attempt(mu, th, w, att >= 62, lam);
#endif
  };
  // ... and the routing of its outcome: an accepted lambda to the Poisson queue of its size, a rejected
  // attempt to the gamma retries.  One slot computation and one 16-byte store serve the three queues.
  auto gamma_route = [&](bool act, bool ok, float lam, float mu, float th, int cell, int gene, int att, uint32_t pword) {
    const bool small = lam < 10.f;
    const unsigned m_s = __ballot_sync(0xffffffffu, act && ok && small);
    const unsigned m_l = __ballot_sync(0xffffffffu, act && ok && !small);
    const unsigned m_g = __ballot_sync(0xffffffffu, act && !ok);
    const int which = ok ? (small ? MX_Q_SMALL : MX_Q_LARGE) : MX_Q_GAMMA;
    const int at = ok ? (small ? ns + __popc(m_s & lt_mask) : nl + __popc(m_l & lt_mask)) : ng + __popc(m_g & lt_mask);
    const float cf = __int_as_float(cell), gf = __int_as_float(gene);
    // fourth word: the inversion's uniform (small lambda) or PTRS attempt 0
    if (act) (&wq.q[0][0])[which * MX_CAP + at] =
        ok ? make_float4(lam, cf, gf, small ? __uint_as_float(pword) : 0.f) : make_float4(mu, th, cf, gf);
    if (act && !ok) wq.g_att[at] = (unsigned char)(att + 1);
    ns += __popc(m_s);
    nl += __popc(m_l);
    ng += __popc(m_g);
  };
  // The uniform of a small-lambda inversion: word (gene mod 4) of block 1 of the (cell, gene quad) stream, so
  // that the lane that owns a quad forms the four of them with ONE Philox block; a gamma retry forms its own.
  auto poisson_block = [&](int cell, uint32_t quad) -> uint4 {
    const int64_t gcell = cell0 + cell;
    return philox_s(key0, key1, quad, (uint32_t)gcell, (TAG_QUAD << 16) | (uint32_t)((uint64_t)gcell >> 32), 1u);
  };
  auto poisson_word = [&](int cell, int gene) -> uint32_t {
    const uint4 b = poisson_block(cell, (uint32_t)gene >> 2);
    const uint32_t j4 = (uint32_t)gene & 3u;
    return j4 == 0 ? b.x : j4 == 1 ? b.y : j4 == 2 ? b.z : b.w;
  };
  auto drain_gamma = [&](int first, int cnt) {
    const bool act = lane < cnt;
    float4 g = make_float4(1.f, 1.f, 0.f, 0.f);
    int att = 1;
    if (act) { g = wq.q[MX_Q_GAMMA][first + lane]; att = wq.g_att[first + lane]; }
    __syncwarp();                                   // every entry is read before its slot is reused
    const int cell = __float_as_int(g.z), gene = __float_as_int(g.w);
    float lam;
    const bool ok = gamma_draw(g.x, g.y, cell, gene, att, lam);
    gamma_route(act, ok, lam, g.x, g.y, cell, gene, att, poisson_word(cell, gene));
    __syncwarp();
  };
  // Poisson(lambda), lambda < 10, by inversion with the uniform that came with the entry: no Philox here
  auto drain_small = [&](int first, int cnt) {
    const bool act = lane < cnt;
    const float4 e = act ? wq.q[MX_Q_SMALL][first + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float lam = e.x;
    const int2 cg = make_int2(__float_as_int(e.y), __float_as_int(e.z));
    const uint32_t word = __float_as_uint(e.w);      // the uniform came with the entry (poisson_word)
    const float u = u01(word);
    float t = exp_fast(-lam);                          // t_k = lambda^k e^-lambda = P(k) k!
    float d = t - u;                                 // cdf(0) - u
    int cn = (int)(__float_as_uint(d) >> 31);        // X = #{k : u > cdf(k)}
#pragma unroll
    for (int k = 1; k <= MX_PS_TERMS; ++k) {
      t *= lam;
      d = fmaf(t, inv_fact[k], d);
      cn += (int)(__float_as_uint(d) >> 31);
    }
    const bool far = word >= 0xFFF00000u;            // decided in fp64 below: not waited for here
    if (far) d = 1.f;
    if (__any_sync(0xffffffffu, d < 0.f)) {          // lambda close to 10 and a large uniform: P(k) form from here
      float pp = t * inv_fact[MX_PS_TERMS];
      for (int k = MX_PS_TERMS; k < 96; k += 4) {
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          pp *= div_fast(lam, (float)(k + s4 + 1));
          d += pp;
          cn += (int)(__float_as_uint(d) >> 31);
        }
        if (!__any_sync(0xffffffffu, d < 0.f)) break;
      }
    }
    if (act && far) {
      // top 2^-12 of the uniforms: fp32 cannot resolve the cdf next to 1 (and its uniform has 24 bits), so
      // the far tail is inverted with 32 more random bits (block 1 of the count's own stream) against a
      // cdf accumulated in fp64
      const int64_t gcell = cell0 + cg.x;
      const uint4 w = philox_s(key0, key1, (uint32_t)cg.y, (uint32_t)gcell,
                               (TAG_COUNT << 16) | (uint32_t)((uint64_t)gcell >> 32), 1u);
      const double u64 = ((double)word + ((double)w.x + 0.5) * 2.3283064365386963e-10) * 2.3283064365386963e-10;
      const double dl = (double)lam;
      double p = exp(-dl), cdf = p, kk = 0.0;
      while (u64 > cdf && kk < 1024.0) { kk += 1.0; p *= dl / kk; cdf += p; }
      cn = (int)kk;
    }
    if (act) write_count(cg.x, cg.y, (float)cn);
    __syncwarp();                                   // every entry is read before its slot is reused
  };
  // Poisson(lambda), lambda >= 10: two PTRS trials from Philox block 2 att + 1; rejected entries come back
  auto drain_large = [&](int first, int cnt) {
    const bool act = lane < cnt;
    const float4 e = act ? wq.q[MX_Q_LARGE][first + lane] : make_float4(100.f, 0.f, 0.f, 0.f);
    __syncwarp();                                   // every entry is read before its slot is reused
    const float lam = e.x;
    const int cell = __float_as_int(e.y), gene = __float_as_int(e.z), att = __float_as_int(e.w);
    const int64_t gcell = cell0 + cell;
#if MX_PTRS_ONE_TRIAL
    // trial number att: Philox block 2 (att / 2) + 1, words (x, y) for an even and (z, w) for an odd trial -
    // the same trials in the same order as two per visit, but a rejected entry goes back to the queue at
    // once, so the second trial's code never runs for the few lanes that need it
    const uint4 w4 = philox_s(key0, key1, (uint32_t)gene, (uint32_t)gcell,
                              (TAG_COUNT << 16) | (uint32_t)((uint64_t)gcell >> 32), ((uint32_t)att | 1u));
    const bool odd = (att & 1) != 0;
    const uint4 w = make_uint4(odd ? w4.z : w4.x, odd ? w4.w : w4.y, 0u, 0u);
    constexpr int kForce = 125;
#else
    const uint4 w = philox_s(key0, key1, (uint32_t)gene, (uint32_t)gcell,
                             (TAG_COUNT << 16) | (uint32_t)((uint64_t)gcell >> 32), 2u * (uint32_t)att + 1u);
    constexpr int kForce = 62;
#endif
    float k;
    bool ok = true;
    if (lam < 1.6e7f) {
      const float slam = lam * rsqrt_fast(lam), loglam = log_fast(lam);
      const float b = fmaf(2.53f, slam, 0.931f);
      const float a = fmaf(0.02483f, b, -0.059f);
      const float log_inv_alpha = log_fast(1.1239f + div_fast(1.1328f, b - 3.4f));
      const float vr = 0.9277f - div_fast(3.6224f, b - 2.0f);
      ok = ptrs_trial(lam, loglam, a, b, log_inv_alpha, vr, u01(w.x) - 0.5f, u01(w.y), k);
#if !MX_PTRS_ONE_TRIAL
      if (!ok) ok = ptrs_trial(lam, loglam, a, b, log_inv_alpha, vr, u01(w.z) - 0.5f, u01(w.w), k);
#endif
      ok = ok || att >= kForce;                      // never reached in practice
      k = fmaxf(k, 0.f);
    } else {                                         // beyond 2^24: normal limit (TV error < 1e-4)
      k = rintf(lam + sqrtf(lam) * sqrtf(-2.0f * log_fast(u01(w.x))) * __cosf(6.2831853072f * u01(w.y) - 3.1415926536f));
    }
    if (act && ok) write_count(cell, gene, k);
    push_large(act && !ok, lam, cell, gene, att + 1);
    __syncwarp();
  };
  // run the queues down to fewer than `limit` entries each (32 while sampling, 1 at the end)
  auto run_queues = [&](int limit) {
    while (ns >= limit || nl >= limit || ng >= limit) {
      while (ns >= limit) { const int c = ns < 32 ? ns : 32; ns -= c; drain_small(ns, c); }
      while (nl >= limit) { const int c = nl < 32 ? nl : 32; nl -= c; drain_large(nl, c); }
      if (ng >= limit) { const int c = ng < 32 ? ng : 32; ng -= c; drain_gamma(ng, c); }
    }
  };

  for (;;) {
    unsigned claimed = 0;
    if (lane == 0) claimed = atomicAdd(&g_sched[sched][0], 1u);
    claimed = __shfl_sync(0xffffffffu, claimed, 0);
    if ((uint64_t)claimed >= n_chunks) break;
    const uint32_t grp = claimed / n_strips;
    const uint32_t strip = claimed - grp * n_strips;
    const uint4 gd = __ldg(groups + grp);                     // first position in `order`, cells, tree row
    const uint32_t pos0 = gd.x;
    const int n_cells = (int)gd.y;
    const bool lane_first = strip * 32u + (uint32_t)lane < Q;     // lanes past the last quad have no work
    const uint32_t quad = min(strip * 32u + (uint32_t)lane, Q - 1u);
    const uint32_t g0 = quad * 4u;
    float m[4], c[4], bm[4];
    bool chunk_clean;
    {
      float al[4];
      const float *mrow = means + (uint64_t)gd.z * G + g0;
      if (VEC) {
        const float4 av = __ldg(reinterpret_cast<const float4 *>(alpha + g0));
        const float4 bv = __ldg(reinterpret_cast<const float4 *>(beta_m1 + g0));
        const float4 mv = ldg_f4_hint(mrow, keep);
        al[0] = av.x; al[1] = av.y; al[2] = av.z; al[3] = av.w;
        bm[0] = bv.x; bm[1] = bv.y; bm[2] = bv.z; bm[3] = bv.w;
        m[0] = mv.x; m[1] = mv.y; m[2] = mv.z; m[3] = mv.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool ok = g0 + j < G;
          al[j] = ok ? alpha[g0 + j] : 0.f;
          bm[j] = ok ? beta_m1[g0 + j] : 1.f;
          m[j] = ok ? mrow[j] : 1.f;
        }
      }
      bool fine = true;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        c[j] = al[j] * m[j];
        fine = fine && (m[j] > 1e-20f) && (m[j] < 1e30f) && (c[j] >= 0.f) && (c[j] < 1e30f) && (bm[j] > 0.f) && (bm[j] < 1e30f);
      }
      chunk_clean = __all_sync(0xffffffffu, fine);
    }
    int32_t meta_cell = 0; float meta_s = 1.f;
    auto load_meta = [&](int first) {                       // cells first .. first+31 of this group
      const int i = first + lane;
      meta_cell = order[pos0 + (uint32_t)(i < n_cells ? i : n_cells - 1)];
      meta_s = scaling[meta_cell];
      // a library size that is not positive and finite is flagged and sampled as NaN (domain error: counts 0)
      if (!(meta_s > 0.f && meta_s < 3.0e38f)) { flag |= PST_FLAG_DOMAIN; meta_s = __int_as_float(0x7fc00000); }
    };
    load_meta(0);
    for (int ci = 0; ci < n_cells; ++ci) {
      const int src = ci & 31;
      const int32_t cell = __shfl_sync(0xffffffffu, meta_cell, src);
      const float s = __shfl_sync(0xffffffffu, meta_s, src);
      if (src == 31) load_meta(ci + 1);
      // the four attempts of the quad first (four independent chains of Philox and MUFU latencies in
      // flight), then their routing
      float mu[4], th[4], lam[4];
      bool go[4], ok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mu[j] = m[j] * s;
        th[j] = fmaf(c[j], s, bm[j]);
        go[j] = lane_first && (VEC || g0 + j < G);
      }
      // mean and theta are positive and finite for every gene of a clean chunk and a library size in range:
      // only the other (cell, strip) pairs check count by count
      if (!(chunk_clean && s > 1e-15f && s < 1e8f)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (!nb_domain_ok(mu[j], th[j])) {               // scipy's "Domain error": flagged, the count is 0
            if (go[j]) {
              flag |= PST_FLAG_DOMAIN;
              X[(uint64_t)(uint32_t)cell * ldx + g0 + j] = 0;
              if constexpr (STATS) atomicAdd(gene_zeros + g0 + j, 1ull);
            }
            go[j] = false;
            mu[j] = 1.f;
            th[j] = 1.f;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) ok[j] = gamma_draw(mu[j], th[j], (int)cell, (int)(g0 + j), 0, lam[j]);
      const uint4 pb = poisson_block((int)cell, quad);
      const uint32_t pw[4] = {pb.x, pb.y, pb.z, pb.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) gamma_route(go[j], ok[j], lam[j], mu[j], th[j], (int)cell, (int)(g0 + j), 0, pw[j]);
      __syncwarp();
      run_queues(32);
    }
  }
  __syncwarp();
  run_queues(1);
  if (flag) atomicOr(flags, flag);
  if (lane == 0) {                                            // the last warp to leave rearms the scheduler words
    const unsigned done = atomicAdd(&g_sched[sched][1], 1u);
    if (done == (unsigned)n_warps - 1u) { g_sched[sched][0] = 0u; g_sched[sched][1] = 0u; __threadfence(); }
  }
}

// Parity hook: the fp32 parameterisation exactly as the draw kernel computes it (same __device__
// functions, same operation order), one element per (M, s, alpha, beta-1).
__global__ void nb_params_f32_kernel(const float *__restrict__ M, const float *__restrict__ scaling,
                                     const float *__restrict__ alpha, const float *__restrict__ beta_m1, int64_t n,
                                     float *__restrict__ out_mu, float *__restrict__ out_theta,
                                     float *__restrict__ out_r, float *__restrict__ out_q, float *__restrict__ out_a,
                                     float *__restrict__ out_log2p0, int32_t *__restrict__ out_route) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float m = M[i], s = scaling[i];
    const float mu = m * s;
    const float th = fmaf(alpha[i] * m, s, beta_m1[i]);
    float q, a, e2;
    nb_inversion_params(mu, th, q, a, e2);
    out_mu[i] = mu;
    out_theta[i] = th;
    out_r[i] = nb_shape(mu, th);
    out_q[i] = q;
    out_a[i] = a;
    out_log2p0[i] = e2;
    const bool s_ok = s > 0.f && s < 3.0e38f;
    const bool inv = s_ok && (s <= nb_inversion_s_max(m, alpha[i], beta_m1[i]));
    out_route[i] = inv ? PST_ROUTE_INVERSION : (s_ok && nb_domain_ok(mu, th)) ? PST_ROUTE_MIXTURE : PST_ROUTE_DOMAIN_ERROR;
  }
}

// The counts the draw kernel listed for the fp64 tail path, one thread each.  tail[0] = number of
// entries, entries (cell, gene, mu, theta) from word 4 on.
// With fused per-gene summaries the draw kernel has counted these elements as zeros: corrected here.
__global__ void tail_fix_kernel(const __grid_constant__ PhiloxKey key, const uint32_t *__restrict__ tail,
                                uint32_t tail_cap, int64_t cell0, int32_t *__restrict__ X, uint32_t ldx,
                                unsigned long long *__restrict__ gene_sum, unsigned long long *__restrict__ gene_sumsq,
                                unsigned long long *__restrict__ gene_zeros) {
  const uint32_t n = min(tail[0], tail_cap);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint4 e = reinterpret_cast<const uint4 *>(tail + 4)[i];
    const int k = invert_tail(__uint_as_float(e.z), __uint_as_float(e.w), key.k0[0], key.k1[0], e.y, cell0 + (int64_t)e.x);
    X[(uint64_t)e.x * ldx + e.y] = k;
    if (gene_sum != nullptr && k > 0) {
      atomicAdd(gene_sum + e.y, (unsigned long long)k);
      atomicAdd(gene_sumsq + e.y, (unsigned long long)k * (unsigned long long)k);
      atomicAdd(gene_zeros + e.y, ~0ull);                   // minus one
    }
  }
}

// ---------------------------------------------------------------------------
// Cells grouped by tree row (counting sort) and cut into groups of at most HY_CHUNK_CELLS cells of one
// row: order[] lists the cells of row 0, then row 1, ...; groups[g] = (first position in order, cells,
// row, 0).  Four small kernels in front of every draw; the layout never changes the counts.
// ---------------------------------------------------------------------------
__global__ void row_histogram_kernel(const int32_t *__restrict__ row_of_cell, int64_t n, int32_t P,
                                     uint32_t *__restrict__ bins, uint32_t *__restrict__ flags) {
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = row_of_cell[i];
    const bool ok = (uint32_t)r < (uint32_t)P;
    bad |= !ok;                                            // flagged (the caller raises), sampled from row 0
    atomicAdd(&bins[ok ? r : 0], 1u);
  }
  if (bad) atomicOr(flags, PST_FLAG_ROW);
}

// exclusive scans over the rows, one CTA: bins[r] -> first position of row r (bins[P] = n),
// gstart[r] -> first group of row r (gstart[P] = number of groups, also written to hdr[1])
__global__ void row_scan_kernel(uint32_t *bins, uint32_t *gstart, int32_t P, uint32_t *hdr) {
  __shared__ uint32_t part[1024], gpart[1024];
  const int t = threadIdx.x;
  const int per = (P + 1023) / 1024;
  const int lo = min(P, t * per), hi = min(P, lo + per);
  uint32_t sum = 0, gsum = 0;
  for (int i = lo; i < hi; ++i) { const uint32_t cnt = bins[i]; sum += cnt; gsum += (cnt + HY_CHUNK_CELLS - 1) / HY_CHUNK_CELLS; }
  part[t] = sum;
  gpart[t] = gsum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const uint32_t v = t >= off ? part[t - off] : 0u, gv = t >= off ? gpart[t - off] : 0u;
    __syncthreads();
    part[t] += v;
    gpart[t] += gv;
    __syncthreads();
  }
  uint32_t run = part[t] - sum, grun = gpart[t] - gsum;
  for (int i = lo; i < hi; ++i) {
    const uint32_t cnt = bins[i];
    bins[i] = run;
    gstart[i] = grun;
    run += cnt;
    grun += (cnt + HY_CHUNK_CELLS - 1) / HY_CHUNK_CELLS;
  }
  if (t == 1023) { bins[P] = part[1023]; gstart[P] = gpart[1023]; hdr[1] = gpart[1023]; }
}

__global__ void group_table_kernel(const uint32_t *__restrict__ bins, const uint32_t *__restrict__ gstart, int32_t P,
                                   uint4 *__restrict__ groups) {
  const uint32_t ng = gstart[P];
  for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += gridDim.x * blockDim.x) {
    int lo = 0, hi = P;                                    // last row r with gstart[r] <= g
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (gstart[mid] <= g) lo = mid; else hi = mid;
    }
    const uint32_t piece = g - gstart[lo];
    const uint32_t pos0 = bins[lo] + piece * HY_CHUNK_CELLS;
    groups[g] = make_uint4(pos0, min((uint32_t)HY_CHUNK_CELLS, bins[lo + 1] - pos0), (uint32_t)lo, 0u);
  }
}

// runs AFTER group_table_kernel: the scatter advances bins[r] from the first to the last position of row r
__global__ void row_scatter_kernel(const int32_t *__restrict__ row_of_cell, int64_t n, int32_t P,
                                   uint32_t *__restrict__ bins, int32_t *__restrict__ order) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t r = row_of_cell[i];
    order[atomicAdd(&bins[(uint32_t)r < (uint32_t)P ? r : 0], 1u)] = (int32_t)i;
  }
}

// ---------------------------------------------------------------------------
// One streaming pass over a count matrix: per-cell totals / zero counts and per-gene sum, sum of
// squares and zero counts (the summary statistics every PROSSTT notebook computes after sampling,
// and the full-size parity check of the sampler).  HBM-read bound: 4 B per count.
// Same strip decomposition as the sampler: a warp owns 32 gene quads (512 B of a row) for a block
// of rows; gene accumulators stay in registers, row partials are warp-reduced and added atomically.
// ---------------------------------------------------------------------------
constexpr int ST_ROWS = 256;                 // rows per chunk (8 groups of 32)
constexpr int ST_BATCH = 8;                  // row loads in flight per thread

template <bool VEC>
__global__ void __launch_bounds__(256, 2)
count_stats_kernel(const int32_t *__restrict__ X, int64_t n, int64_t G, int64_t ldx,
                   unsigned long long *__restrict__ cell_total, unsigned int *__restrict__ cell_zeros,
                   unsigned long long *__restrict__ gene_sum, unsigned long long *__restrict__ gene_sumsq,
                   unsigned long long *__restrict__ gene_zeros) {
  typedef unsigned long long ull;
  const int lane = threadIdx.x & 31;
  const int64_t Q = (G + 3) / 4;
  const int64_t n_strips = (Q + 31) / 32;
  const int64_t n_chunks = ((n + ST_ROWS - 1) / ST_ROWS) * n_strips;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t chunk = warp0; chunk < n_chunks; chunk += n_warps) {
    const int64_t rgroup = chunk / n_strips;
    const int64_t quad = (chunk - rgroup * n_strips) * 32 + lane;
    const bool lane_ok = quad < Q;
    const int64_t g0 = (lane_ok ? quad : 0) * 4;
    const int64_t r_lo = rgroup * ST_ROWS;
    const int64_t r_hi = min(n, r_lo + ST_ROWS);
    ull s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    unsigned int z[4] = {0, 0, 0, 0};
    for (int64_t rb = r_lo; rb < r_hi; rb += 16) {
      // this lane's partial of 16 rows: total in bits 0..39, zero count in bits 40..47
      ull part[16];
#pragma unroll
      for (int b0 = 0; b0 < 16; b0 += ST_BATCH) {
        int4 q[ST_BATCH];
#pragma unroll
        for (int i = 0; i < ST_BATCH; ++i) {
          const int64_t r = rb + b0 + i;
          q[i] = make_int4(0, 0, 0, 0);
          if (lane_ok && r < r_hi) {
            if (VEC) {
              q[i] = __ldcs(reinterpret_cast<const int4 *>(X + r * ldx + g0));
            } else {
              const int32_t *src = X + r * ldx + g0;
              q[i].x = src[0];
              if (g0 + 1 < G) q[i].y = src[1];
              if (g0 + 2 < G) q[i].z = src[2];
              if (g0 + 3 < G) q[i].w = src[3];
            }
          }
        }
#pragma unroll
        for (int i = 0; i < ST_BATCH; ++i) {
          const bool row_in = lane_ok && (rb + b0 + i < r_hi);
          const unsigned int v[4] = {(unsigned int)q[i].x, (unsigned int)q[i].y, (unsigned int)q[i].z,
                                     (unsigned int)q[i].w};
          ull rt = 0;
          unsigned int rz = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool in = row_in && (VEC || g0 + j < G);
            s[j] += v[j];
            ss[j] += (ull)v[j] * v[j];
            const unsigned int zero = (in && v[j] == 0u) ? 1u : 0u;
            z[j] += zero;
            rt += v[j];
            rz += zero;
          }
          part[b0 + i] = rt | ((ull)rz << 40);
        }
      }
      // transpose-reduce: 15 exchanges leave the sum over each 16-lane half of row (rb + lane%16)
      // in part[0]; one more exchange adds the two halves
#pragma unroll
      for (int half = 8; half > 0; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const ull send = upper ? part[i] : part[i + half];
          const ull keep = upper ? part[i + half] : part[i];
          const unsigned int lo = __shfl_xor_sync(0xffffffffu, (unsigned int)send, half);
          const unsigned int hi = __shfl_xor_sync(0xffffffffu, (unsigned int)(send >> 32), half);
          part[i] = keep + (((ull)hi << 32) | lo);
        }
      }
      {
        const unsigned int lo = __shfl_xor_sync(0xffffffffu, (unsigned int)part[0], 16);
        const unsigned int hi = __shfl_xor_sync(0xffffffffu, (unsigned int)(part[0] >> 32), 16);
        part[0] += ((ull)hi << 32) | lo;
      }
      const int64_t r = rb + (lane & 15);
      if (lane < 16 && r < r_hi) {
        if (cell_total) atomicAdd(cell_total + r, part[0] & 0xFFFFFFFFFFull);
        if (cell_zeros) atomicAdd(cell_zeros + r, (unsigned int)(part[0] >> 40));
      }
    }
    if (lane_ok) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (VEC || g0 + j < G) {
          if (gene_sum) atomicAdd(gene_sum + g0 + j, s[j]);
          if (gene_sumsq) atomicAdd(gene_sumsq + g0 + j, ss[j]);
          if (gene_zeros) atomicAdd(gene_zeros + g0 + j, (ull)z[j]);
        }
      }
    }
  }
}

}  // namespace pst

using namespace pst;

extern "C" int pst_count_stats(const int32_t *X, int64_t n, int64_t G, int64_t ldx, uint64_t *cell_total,
                               uint32_t *cell_zeros, uint64_t *gene_sum, uint64_t *gene_sumsq,
                               uint64_t *gene_zeros, void *stream) {
  const char *fn = "pst_count_stats";
  PST_REQUIRE(n >= 0 && G >= 0 && ldx >= G, fn, "need n, G >= 0 and ldx >= G");
  if (n == 0 || G == 0) return 0;
  PST_REQUIRE(X, fn, "null pointer");
  const bool vec = (G % 4 == 0) && (ldx % 4 == 0) && ((uintptr_t)X % 16 == 0);
  const int64_t n_chunks = ((n + ST_ROWS - 1) / ST_ROWS) * ((((G + 3) / 4) + 31) / 32);
  const int64_t need = (n_chunks + 7) / 8;                        // 8 warps per CTA
  const unsigned blocks = (unsigned)std::min<int64_t>(need, (int64_t)num_sm() * 2);   // 2 CTAs of 8 warps per SM
  typedef unsigned long long ull;
  if (vec)
    count_stats_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        X, n, G, ldx, (ull *)cell_total, cell_zeros, (ull *)gene_sum, (ull *)gene_sumsq, (ull *)gene_zeros);
  else
    count_stats_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        X, n, G, ldx, (ull *)cell_total, cell_zeros, (ull *)gene_sum, (ull *)gene_sumsq, (ull *)gene_zeros);
  return check_launch(fn);
}


extern "C" int pst_nb_params_f32(const float *M, const float *scaling, const float *alpha, const float *beta_m1,
                                 int64_t n, float *out_mu, float *out_theta, float *out_r, float *out_q, float *out_a,
                                 float *out_log2p0, int32_t *out_route, void *stream) {
  const char *fn = "pst_nb_params_f32";
  PST_REQUIRE(n >= 0, fn, "negative size");
  if (n == 0) return 0;
  PST_REQUIRE(M && scaling && alpha && beta_m1 && out_mu && out_theta && out_r && out_q && out_a && out_log2p0 &&
              out_route, fn, "null pointer");
  const unsigned g = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)num_sm() * 8);
  nb_params_f32_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(M, scaling, alpha, beta_m1, n, out_mu, out_theta, out_r,
                                                            out_q, out_a, out_log2p0, out_route);
  return check_launch(fn);
}

// scratch layout of pst_draw_counts (uint32 words): [0] tail entries appended, [1] number of groups,
// [2..3] unused | tail list, 4 words per entry | order[n] | bins[P+1] | gstart[P+1] | groups, 4 words each
namespace {
struct DrawLayout { int64_t tail_cap, order, bins, gstart, groups, max_groups, words; };
DrawLayout draw_layout(int64_t n, int64_t G, int64_t P) {
  DrawLayout L;
  // twice the expected 2^-14 n G listed counts plus 1024: a list that does not fit would be a > 30 sigma event
  const double expect = (double)n * (double)G / 16384.0;
  L.tail_cap = std::min<int64_t>((int64_t)(2.0 * expect) + 1024, (int64_t)0xfffffff0);
  L.order = 4 + 4 * L.tail_cap;
  L.bins = L.order + n;
  L.gstart = L.bins + P + 1;
  L.groups = (L.gstart + P + 1 + 3) & ~(int64_t)3;          // uint4 entries: 16-byte aligned
  L.max_groups = n / HY_CHUNK_CELLS + P + 1;
  L.words = L.groups + 4 * L.max_groups;
  return L;
}
}  // namespace

extern "C" int64_t pst_draw_scratch_words(int64_t n, int64_t G, int64_t P) {
  if (n <= 0 || G <= 0 || P <= 0) return 4;
  return draw_layout(n, G, P).words;
}

extern "C" int pst_draw_scratch_layout(int64_t n, int64_t G, int64_t P, int64_t *h_out) {
  if (n <= 0 || G <= 0 || P <= 0 || !h_out) return -1;
  const DrawLayout L = draw_layout(n, G, P);
  h_out[0] = L.tail_cap; h_out[1] = L.order; h_out[2] = L.bins; h_out[3] = L.gstart; h_out[4] = L.groups;
  h_out[5] = L.max_groups;
  return 0;
}

extern "C" int pst_draw_counts(const float *means, int64_t P, int64_t G, const int32_t *row_of_cell,
                               const float *scaling, const float *alpha, const float *beta_m1,
                               uint64_t seed, int64_t cell0, int64_t n, int32_t *X, int64_t ldx,
                               uint32_t *flags, int32_t sampler, uint32_t *scratch, int64_t scratch_words,
                               uint64_t *gene_sum, uint64_t *gene_sumsq, uint64_t *gene_zeros, void *stream) {
  const char *fn = "pst_draw_counts";
  const bool stats = gene_sum || gene_sumsq || gene_zeros;
  PST_REQUIRE(!stats || (gene_sum && gene_sumsq && gene_zeros), fn,
              "the fused per-gene summaries come as a set: pass all three arrays or none");
  PST_REQUIRE(P >= 0 && G >= 0 && n >= 0 && cell0 >= 0, fn, "negative size");
  PST_REQUIRE(ldx >= G, fn, "ldx < G");
  PST_REQUIRE(G < (int64_t)1 << 31 && P < (int64_t)1 << 31 && ldx < (int64_t)1 << 32, fn,
              "G and P must be below 2^31 and ldx below 2^32");
  PST_REQUIRE(n < (int64_t)1 << 31, fn, "at most 2^31-1 cells per call (shard or chunk the cells)");
  PST_REQUIRE(cell0 + n < (int64_t)1 << 48, fn, "cell index must be below 2^48");
  PST_REQUIRE(sampler == PST_SAMPLER_GAMMA_POISSON || sampler == PST_SAMPLER_HYBRID, fn, "unknown sampler");
  if (n == 0 || G == 0) return 0;
  PST_REQUIRE(P > 0, fn, "the means table has no rows");
  PST_REQUIRE(means && row_of_cell && scaling && alpha && beta_m1 && X && flags, fn, "null pointer");
  const DrawLayout L = draw_layout(n, G, P);
  PST_REQUIRE(scratch && scratch_words >= L.words, fn, "scratch must hold pst_draw_scratch_words(n, G, P) words");
  PST_REQUIRE((uintptr_t)scratch % 16 == 0, fn, "scratch must be 16-byte aligned");
  const int64_t Q = (G + 3) / 4;
  const bool vec = (G % 4 == 0) && (ldx % 4 == 0) && ((uintptr_t)means % 16 == 0) &&
                   ((uintptr_t)alpha % 16 == 0) && ((uintptr_t)beta_m1 % 16 == 0) &&
                   ((uintptr_t)X % 16 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_strips = (Q + 31) / 32;
  PST_REQUIRE(L.max_groups * n_strips < ((int64_t)1 << 32) - 65536, fn, "too many work chunks in one call (chunk the cells)");
  const int slot = sched_slot(stream);
  PST_REQUIRE(slot >= 0, fn, "more than 4096 distinct streams have called pst_draw_counts on this device");

  // ---- cells grouped by tree row, cut into groups of one row
  uint32_t *bins = scratch + L.bins, *gstart = scratch + L.gstart;
  int32_t *order = reinterpret_cast<int32_t *>(scratch + L.order);
  uint4 *groups = reinterpret_cast<uint4 *>(scratch + L.groups);
  cudaError_t e = cudaMemsetAsync(scratch, 0, 16, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(bins, 0, sizeof(uint32_t) * (size_t)(P + 1), st);
  if (e != cudaSuccess) { cudaGetLastError(); return pst::fail_arg(fn, "memset failed"); }
  const unsigned gg = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)num_sm() * 8);
  row_histogram_kernel<<<gg, 256, 0, st>>>(row_of_cell, n, (int32_t)P, bins, flags);
  int rc = check_launch(fn);
  if (rc) return rc;
  row_scan_kernel<<<1, 1024, 0, st>>>(bins, gstart, (int32_t)P, scratch);
  rc = check_launch(fn);
  if (rc) return rc;
  const unsigned gt = (unsigned)std::min<int64_t>((L.max_groups + 255) / 256, (int64_t)num_sm() * 8);
  group_table_kernel<<<gt, 256, 0, st>>>(bins, gstart, (int32_t)P, groups);
  rc = check_launch(fn);
  if (rc) return rc;
  row_scatter_kernel<<<gg, 256, 0, st>>>(row_of_cell, n, (int32_t)P, bins, order);
  rc = check_launch(fn);
  if (rc) return rc;

  // ---- the draw: persistent CTAs of 4 warps, one wave; never more warps than chunks can exist
  const int64_t need = (L.max_groups * n_strips + HY_WARPS - 1) / HY_WARPS;
  const int64_t cap = (int64_t)num_sm() * (stats ? HY_STATS_CTAS : HY_MIN_CTAS);
  const unsigned hb = (unsigned)(need < cap ? need : cap);
  const uint32_t tail_cap = (uint32_t)L.tail_cap;
  typedef unsigned long long ull;
#define PST_LAUNCH_DRAW(V, ST)                                                                                      \
    draw_counts_kernel<HY_KFIX, V, ST><<<hb, HY_THREADS, 0, st>>>(                                                  \
        PhiloxKey(seed), means, (uint32_t)G, (uint32_t)Q, scaling, alpha, beta_m1, cell0, X, (uint32_t)ldx, flags,  \
        scratch, order, groups, slot, scratch, tail_cap, (ull *)gene_sum, (ull *)gene_sumsq, (ull *)gene_zeros)
  if (sampler == PST_SAMPLER_GAMMA_POISSON) {            // every count through the mixture pipeline
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const int64_t mcap = (int64_t)num_sm() * MX_MIN_CTAS;
    const unsigned mb = (unsigned)(need < mcap ? need : mcap);
#define PST_LAUNCH_MIX(V, ST)                                                                                       \
    draw_counts_mixture_kernel<V, ST><<<mb, HY_THREADS, 0, st>>>(                                                   \
        k0, k1, means, (uint32_t)G, (uint32_t)Q, scaling, alpha, beta_m1, cell0, X, (uint32_t)ldx, flags, scratch,  \
        order, groups, slot, (ull *)gene_sum, (ull *)gene_sumsq, (ull *)gene_zeros)
    if (vec && stats) PST_LAUNCH_MIX(true, true);
    else if (vec) PST_LAUNCH_MIX(true, false);
    else if (stats) PST_LAUNCH_MIX(false, true);
    else PST_LAUNCH_MIX(false, false);
#undef PST_LAUNCH_MIX
    return check_launch(fn);
  }
  if (vec && stats) PST_LAUNCH_DRAW(true, true);
  else if (vec) PST_LAUNCH_DRAW(true, false);
  else if (stats) PST_LAUNCH_DRAW(false, true);
  else PST_LAUNCH_DRAW(false, false);
#undef PST_LAUNCH_DRAW
  rc = check_launch(fn);
  if (rc) return rc;
  // the listed counts (2^-14 of the inverted ones): one thread each, grid-stride beyond the expectation
  const int64_t expect = (int64_t)((double)n * (double)G / 16384.0) + 256;
  const unsigned tb = (unsigned)std::min<int64_t>((expect + 127) / 128, (int64_t)num_sm() * 16);
  tail_fix_kernel<<<tb, 128, 0, st>>>(PhiloxKey(seed), scratch, tail_cap, cell0, X, (uint32_t)ldx, (ull *)gene_sum,
                                      (ull *)gene_sumsq, (ull *)gene_zeros);
  return check_launch(fn);
}
