// Lineage stage: walk draws, momentum-walk scan, parent carry, rel = W.H fused with
// exp(.)*gene_scale and the per-gene maximum, sibling Pearson check.
// Reference behaviour: prosstt/simulation.py:21-124, 215-286; prosstt/sim_utils.py:129-168,
// 190-213, 406-426, 611-640; prosstt/tree.py:166-183.  fp64 throughout (parity 1e-6 rel).
#include "pst_common.cuh"

namespace pst {

// ---------------------------------------------------------------------------
// draws of simulation.diffusion, counter = (t, k | attempt<<16, tag, branch_id)
// ---------------------------------------------------------------------------
__global__ void walk_draws_kernel(PhiloxKey key, int nb, int K, const int32_t *__restrict__ branch_id,
                                  const int32_t *__restrict__ attempt, const int32_t *__restrict__ T,
                                  const int64_t *__restrict__ eps_off, double *__restrict__ u0,
                                  double *__restrict__ v0, double *__restrict__ eta,
                                  double *__restrict__ eps) {
  const int j = blockIdx.y;            // branch slot
  const int k = blockIdx.z;            // program
  const int Tj = T[j];
  const uint32_t c1 = (uint32_t)k | ((uint32_t)attempt[j] << 16);
  const uint32_t c3 = (uint32_t)branch_id[j];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) {
    // simulation.py:107-112: U(0,1.5), N(0,0.2), U(0,1)
    uint4 r = philox(key, 0u, c1, TAG_WALK_U0, c3);
    u0[j * K + k] = 1.5 * u53(r.x, r.y);
    eta[j * K + k] = u53(r.z, r.w);
    v0[j * K + k] = 0.2 * normal_f64(philox(key, 0u, c1, TAG_WALK_V0, c3));
  }
  if (t < Tj - 1) {
    // simulation.py:111,117: eps_t ~ N(0, 2/steps)
    const double z = normal_f64(philox(key, (uint32_t)t, c1, TAG_WALK_EPS, c3));
    eps[eps_off[j] + (int64_t)k * (Tj - 1) + t] = (2.0 / (double)Tj) * z;
  }
}

// ---------------------------------------------------------------------------
// momentum walk as a warp scan; one warp per (branch slot, program)
//   v[t+1] = eta v[t] + eps[t]   -> affine-map scan: b_i = sum_j eta^(i-j) eps_j
//   walk[t+1] = walk[t] + v[t]   -> prefix sum of v
// ---------------------------------------------------------------------------
__global__ void walk_scan_kernel(int nb, int K, const int32_t *__restrict__ row_base,
                                 const int32_t *__restrict__ T, const int64_t *__restrict__ eps_off,
                                 const double *__restrict__ u0, const double *__restrict__ v0,
                                 const double *__restrict__ eta, const double *__restrict__ eps,
                                 double *__restrict__ W) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nb * K) return;
  const int j = warp / K, k = warp - j * K;
  const int Tj = T[j];
  const double e = eta[warp];
  double epow[6];                       // eta^(1,2,4,8,16,32)
  epow[0] = e;
#pragma unroll
  for (int i = 1; i < 6; ++i) epow[i] = epow[i - 1] * epow[i - 1];
  // eta^(lane+1) by binary powering
  double elane = 1.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) if ((lane + 1) >> i & 1) elane *= epow[i];

  const double *ep = eps + eps_off[j] + (int64_t)k * (Tj - 1);
  double *out = W + (int64_t)row_base[j] * K + k;
  double walk = log(u0[warp]);          // walk[base]
  double vel = v0[warp];                // v[base]
  if (lane == 0) out[0] = walk;
  for (int base = 0; base < Tj - 1; base += 32) {
    const int t = base + lane;          // this lane owns step t -> t+1
    double b = (t < Tj - 1) ? ep[t] : 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const double up = __shfl_up_sync(0xffffffffu, b, 1 << i);
      if (lane >= (1 << i)) b = fma(epow[i], up, b);
    }
    const double vnew = fma(elane, vel, b);           // v[base+lane+1]
    // x_0 = v[base], x_i = v[base+i]  -> inclusive sum gives walk[base+i+1]-walk[base]
    double x = __shfl_up_sync(0xffffffffu, vnew, 1);
    if (lane == 0) x = vel;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const double up = __shfl_up_sync(0xffffffffu, x, 1 << i);
      if (lane >= (1 << i)) x += up;
    }
    const double wnew = walk + x;                     // walk[base+lane+1]
    if (t < Tj - 1) out[(int64_t)(t + 1) * K] = wnew;
    walk = __shfl_sync(0xffffffffu, wnew, 31);
    vel = __shfl_sync(0xffffffffu, vnew, 31);
  }
}

// ---------------------------------------------------------------------------
// parent carry, one block per program column, branches in topological order
// ---------------------------------------------------------------------------
__global__ void walk_carry_kernel(int n_order, int K, const int32_t *__restrict__ order_row_base,
                                  const int32_t *__restrict__ order_T,
                                  const int32_t *__restrict__ order_parent_last, double *W) {
  const int k = blockIdx.x;
  __shared__ double dif;
  for (int i = 0; i < n_order; ++i) {
    const int pl = order_parent_last[i];
    if (pl < 0) continue;                              // root: nothing to adjust (block-uniform)
    double *col = W + (int64_t)order_row_base[i] * K + k;
    if (threadIdx.x == 0) dif = col[0] - W[(int64_t)pl * K + k];   // sim_utils.py:140
    __syncthreads();
    const double d = dif;
    for (int t = threadIdx.x; t < order_T[i]; t += blockDim.x) col[(int64_t)t * K] -= d;  // :141
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// rel = W.H (+ exp * gene_scale, + per-gene max).  Skinny contraction, K ~ 10:
// a block owns 256 genes; the W rows of a tile sit in shared memory and are
// broadcast; each thread keeps RT accumulators so one H load feeds RT FMAs.
// ---------------------------------------------------------------------------
constexpr double kMinPositiveMean = 1e-30;
constexpr int RM_THREADS = 256;
constexpr int RM_RT = 8;               // rows per register tile

__global__ void __launch_bounds__(RM_THREADS)
rel_means_kernel(const double *__restrict__ W, const double *__restrict__ H,
                 const double *__restrict__ gene_scale, int64_t row0, int64_t nrows, int K,
                 int64_t G, int64_t rows_per_block, double *__restrict__ out_rel,
                 double *__restrict__ out_m64, float *__restrict__ out_m32,
                 double *__restrict__ out_colmax) {
  extern __shared__ double wtile[];     // [RM_RT][K]
  const int64_t g = (int64_t)blockIdx.x * RM_THREADS + threadIdx.x;
  const int64_t rbeg = row0 + (int64_t)blockIdx.y * rows_per_block;
  const int64_t rend = min(row0 + nrows, rbeg + rows_per_block);
  const double gs = (gene_scale != nullptr && g < G) ? gene_scale[g] : 1.0;
  double cmax = -INFINITY;
  for (int64_t r = rbeg; r < rend; r += RM_RT) {
    const int nr = (int)min((int64_t)RM_RT, rend - r);
    __syncthreads();
    for (int i = threadIdx.x; i < RM_RT * K; i += RM_THREADS)
      wtile[i] = (i < nr * K) ? W[r * K + i] : 0.0;
    __syncthreads();
    if (g < G) {
      double acc[RM_RT];
#pragma unroll
      for (int i = 0; i < RM_RT; ++i) acc[i] = 0.0;
      for (int k = 0; k < K; ++k) {
        const double h = H[(int64_t)k * G + g];
#pragma unroll
        for (int i = 0; i < RM_RT; ++i) acc[i] = fma(wtile[i * K + k], h, acc[i]);
      }
#pragma unroll
      for (int i = 0; i < RM_RT; ++i) {
        if (i < nr) {
          const int64_t o = (r + i) * G + g;
          cmax = fmax(cmax, acc[i]);
          if (out_rel) out_rel[o] = acc[i];
          if (out_m64 || out_m32) {
            const double m = exp(acc[i]) * gs;
            if (out_m64) out_m64[o] = m;
            // a positive mean below the fp32 range would read as 0 (= scipy's domain error):
            // keep it positive; the count is 0 with probability 1 - 1e-30 either way
            if (out_m32) out_m32[o] = (m > 0.0 && m < kMinPositiveMean) ? (float)kMinPositiveMean : (float)m;
          }
        }
      }
    }
  }
  if (out_colmax && g < G && rend > rbeg) atomic_max_f64(out_colmax + g, cmax);
}

// ---------------------------------------------------------------------------
// genes with Pearson r < 0 between two (n x G) slabs; one thread per gene, rows
// coalesced across the warp.  Centre first, like scipy.stats.pearsonr.
// ---------------------------------------------------------------------------
__global__ void pearson_kernel(const double *__restrict__ A, const double *__restrict__ B,
                               int64_t n, int64_t G, int32_t *out_count) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int neg = 0;
  if (g < G) {
    double sa = 0, sb = 0;
    for (int64_t t = 0; t < n; ++t) { sa += A[t * G + g]; sb += B[t * G + g]; }
    const double ma = sa / (double)n, mb = sb / (double)n;
    double sab = 0, saa = 0, sbb = 0;
    for (int64_t t = 0; t < n; ++t) {
      const double a = A[t * G + g] - ma, b = B[t * G + g] - mb;
      sab = fma(a, b, sab); saa = fma(a, a, saa); sbb = fma(b, b, sbb);
    }
    neg = (saa > 0.0 && sbb > 0.0 && sab < 0.0) ? 1 : 0;   // NaN r (constant column) is not < 0
  }
  const unsigned m = __ballot_sync(0xffffffffu, neg);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out_count, __popc(m));
}

// ---------------------------------------------------------------------------
// Acceptance tests of simulate_lineage for a whole batch of candidate branches in ONE launch
// (simulation.py:270-272): item i < n_slots -> maximum of the candidate's (T x G) rel rows (cutoff
// test); item n_slots + p -> number of genes with Pearson r < 0 between two row blocks (sibling
// divergence, same arithmetic as pearson_kernel).  blockIdx.y = item, blockIdx.x = 128-gene block.
// ---------------------------------------------------------------------------
__global__ void lineage_checks_kernel(const double *__restrict__ rel, int64_t G, int n_slots,
                                      const int64_t *__restrict__ slot_row, const int32_t *__restrict__ slot_T,
                                      const int64_t *__restrict__ pair_a, const int64_t *__restrict__ pair_b,
                                      const int32_t *__restrict__ pair_n, double *__restrict__ slot_max,
                                      int32_t *__restrict__ pair_neg) {
  const int item = blockIdx.y;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item < n_slots) {
    const double *A = rel + slot_row[item] * G;
    const int64_t n = slot_T[item];
    double m = -INFINITY;
    if (g < G)
      for (int64_t t = 0; t < n; ++t) m = fmax(m, A[t * G + g]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0 && m > -INFINITY) atomic_max_f64(slot_max + item, m);
    return;
  }
  const int p = item - n_slots;
  const double *A = rel + pair_a[p] * G, *B = rel + pair_b[p] * G;
  const int64_t n = pair_n[p];
  int neg = 0;
  if (g < G) {
    double sa = 0, sb = 0;
    for (int64_t t = 0; t < n; ++t) { sa += A[t * G + g]; sb += B[t * G + g]; }
    const double ma = sa / (double)n, mb = sb / (double)n;
    double sab = 0, saa = 0, sbb = 0;
    for (int64_t t = 0; t < n; ++t) {
      const double a = A[t * G + g] - ma, b = B[t * G + g] - mb;
      sab = fma(a, b, sab); saa = fma(a, a, saa); sbb = fma(b, b, sbb);
    }
    neg = (saa > 0.0 && sbb > 0.0 && sab < 0.0) ? 1 : 0;   // NaN r (constant column) is not < 0
  }
  const unsigned m = __ballot_sync(0xffffffffu, neg);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(pair_neg + p, __popc(m));
}

__global__ void f64_to_f32_kernel(const double *__restrict__ in, int64_t n, double min_positive,
                                  float *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double v = in[i];
    out[i] = (v > 0.0 && v < min_positive) ? (float)min_positive : (float)v;
  }
}

}  // namespace pst

using namespace pst;

extern "C" int pst_walk_draws(uint64_t seed, int32_t nb, int32_t K, int32_t max_T,
                              const int32_t *branch_id, const int32_t *attempt, const int32_t *T,
                              const int64_t *eps_off, double *u0, double *v0, double *eta,
                              double *eps, void *stream) {
  const char *fn = "pst_walk_draws";
  PST_REQUIRE(nb >= 0 && K >= 0 && max_T >= 0, fn, "nb, K, max_T >= 0 required");
  if (nb == 0 || K == 0) return 0;
  PST_REQUIRE(nb <= 65535 && K <= 65535, fn, "at most 65535 branches and programs per call");
  PST_REQUIRE(branch_id && attempt && T && eps_off && u0 && v0 && eta && eps, fn, "null pointer");
  dim3 grid((unsigned)((max_T > 1 ? max_T - 1 : 1) + 255) / 256, nb, K);
  walk_draws_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(PhiloxKey(seed), nb, K, branch_id, attempt, T,
                                                            eps_off, u0, v0, eta, eps);
  return check_launch(fn);
}

extern "C" int pst_walk_scan(int32_t nb, int32_t K, const int32_t *row_base, const int32_t *T,
                             const int64_t *eps_off, const double *u0, const double *v0,
                             const double *eta, const double *eps, double *W, void *stream) {
  const char *fn = "pst_walk_scan";
  PST_REQUIRE(nb >= 0 && K >= 0, fn, "nb, K >= 0 required");
  if (nb == 0 || K == 0) return 0;
  PST_REQUIRE(row_base && T && eps_off && u0 && v0 && eta && eps && W, fn, "null pointer");
  const int64_t warps = (int64_t)nb * K;
  const int threads = 128;
  const int64_t blocks = (warps * 32 + threads - 1) / threads;
  walk_scan_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(nb, K, row_base, T, eps_off, u0,
                                                                           v0, eta, eps, W);
  return check_launch(fn);
}

extern "C" int pst_walk_carry(int32_t n_order, int32_t K, const int32_t *order_row_base,
                              const int32_t *order_T, const int32_t *order_parent_last, double *W,
                              void *stream) {
  const char *fn = "pst_walk_carry";
  PST_REQUIRE(n_order >= 0 && K >= 0, fn, "n_order, K >= 0 required");
  if (n_order == 0 || K == 0) return 0;
  PST_REQUIRE(order_row_base && order_T && order_parent_last && W, fn, "null pointer");
  walk_carry_kernel<<<K, 256, 0, (cudaStream_t)stream>>>(n_order, K, order_row_base, order_T,
                                                         order_parent_last, W);
  return check_launch(fn);
}

extern "C" int pst_rel_means(const double *W, const double *H, const double *gene_scale,
                             int64_t row0, int64_t nrows, int32_t K, int64_t G, double *out_rel,
                             double *out_mean64, float *out_mean32, double *out_colmax,
                             void *stream) {
  const char *fn = "pst_rel_means";
  PST_REQUIRE(row0 >= 0 && nrows >= 0 && K >= 0 && G >= 0, fn, "negative size");
  if (nrows == 0 || G == 0) return 0;
  PST_REQUIRE(W || K == 0, fn, "W is null");
  PST_REQUIRE(H || K == 0, fn, "H is null");
  PST_REQUIRE(K <= 1024, fn, "K > 1024 programs not supported");
  const int64_t gx = (G + RM_THREADS - 1) / RM_THREADS;
  // enough row chunks for ~4 CTAs per SM, each a multiple of the register tile
  int64_t want_y = (4 * num_sm() + gx - 1) / gx;
  int64_t rows_per_block = (nrows + want_y - 1) / want_y;
  rows_per_block = ((rows_per_block + RM_RT - 1) / RM_RT) * RM_RT;
  const int64_t gy = (nrows + rows_per_block - 1) / rows_per_block;
  PST_REQUIRE(gy <= 65535, fn, "too many row chunks");
  dim3 grid((unsigned)gx, (unsigned)gy);
  const size_t smem = sizeof(double) * RM_RT * (size_t)(K > 0 ? K : 1);
  rel_means_kernel<<<grid, RM_THREADS, smem, (cudaStream_t)stream>>>(
      W, H, gene_scale, row0, nrows, K, G, rows_per_block, out_rel, out_mean64, out_mean32, out_colmax);
  return check_launch(fn);
}

extern "C" int pst_pearson_anticorr(const double *A, const double *B, int64_t nrows, int64_t G,
                                    int32_t *out_count, void *stream) {
  const char *fn = "pst_pearson_anticorr";
  PST_REQUIRE(nrows >= 0 && G >= 0, fn, "negative size");
  if (G == 0) return 0;
  PST_REQUIRE(A && B && out_count, fn, "null pointer");
  pearson_kernel<<<(unsigned)((G + 127) / 128), 128, 0, (cudaStream_t)stream>>>(A, B, nrows, G, out_count);
  return check_launch(fn);
}

extern "C" int pst_lineage_checks(const double *rel, int64_t G, int32_t n_slots, const int64_t *slot_row,
                                  const int32_t *slot_T, int32_t n_pairs, const int64_t *pair_a,
                                  const int64_t *pair_b, const int32_t *pair_n, double *out_slot_max,
                                  int32_t *out_pair_neg, void *stream) {
  const char *fn = "pst_lineage_checks";
  PST_REQUIRE(G >= 0 && n_slots >= 0 && n_pairs >= 0, fn, "negative size");
  PST_REQUIRE((int64_t)n_slots + n_pairs <= 65535, fn, "at most 65535 candidates + pairs per call");
  if (G == 0 || n_slots + n_pairs == 0) return 0;
  PST_REQUIRE(rel, fn, "null pointer");
  PST_REQUIRE(n_slots == 0 || (slot_row && slot_T && out_slot_max), fn, "null pointer (candidates)");
  PST_REQUIRE(n_pairs == 0 || (pair_a && pair_b && pair_n && out_pair_neg), fn, "null pointer (pairs)");
  const dim3 grid((unsigned)((G + 127) / 128), (unsigned)(n_slots + n_pairs));
  lineage_checks_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(rel, G, n_slots, slot_row, slot_T, pair_a, pair_b,
                                                                pair_n, out_slot_max, out_pair_neg);
  return check_launch(fn);
}

extern "C" int pst_f64_to_f32(const double *in, int64_t n, double min_positive, float *out, void *stream) {
  const char *fn = "pst_f64_to_f32";
  PST_REQUIRE(n >= 0, fn, "negative size");
  if (n == 0) return 0;
  PST_REQUIRE(in && out, fn, "null pointer");
  const int64_t blocks = min((n + 255) / 256, (int64_t)num_sm() * 16);
  f64_to_f32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, n, min_positive, out);
  return check_launch(fn);
}
