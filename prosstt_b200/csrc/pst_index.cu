// Counter-based draws and the samplers' cell -> (pseudotime, branch, packed row) index maps.
// Reference behaviour: prosstt/simulation.py:319-548 (sample_density, sample_pseudotime_series,
// draw_times, sample_whole_tree), prosstt/sim_utils.py:342-403 (pick_branches/pick_branch),
// :473-498 (calc_scalings).  fp64 + integers; maps are bit-exact given the same uniforms/normals.
#include <map>
#include <mutex>
#include <utility>
#include "pst_common.cuh"

namespace pst {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int num_sm() {
  static std::atomic<int> cached[64];                 // per device ordinal; 0 = not queried yet
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return 148; }
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = 148;                                        // B200: 2 dies x 74 SMs
    }
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// pst_draw_counts hands out its work chunks through a device-side counter.  The counters live in a
// library-owned __device__ array (one instance per GPU); every (device, stream) pair gets its own
// slot, so launches on different streams never share one and launches on one stream are ordered.
int sched_slot(void *stream) {
  static std::mutex mu;
  static std::map<std::pair<int, void *>, int> slots;
  static int next_slot[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); dev = 0; }
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_pair(dev, stream);
  const auto it = slots.find(key);
  if (it != slots.end()) return it->second;
  if (next_slot[dev] >= PST_SCHED_SLOTS) return -1;
  const int s = next_slot[dev]++;
  slots[key] = s;
  return s;
}

int fail_arg(const char *fn, const char *what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", fn, what);
  return -1;
}

int check_launch(const char *fn) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", fn, (int)e, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

__device__ __forceinline__ uint4 element_block(const PhiloxKey &key, uint32_t tag, int64_t idx) {
  return philox(key, (uint32_t)idx, (uint32_t)((uint64_t)idx >> 32), tag, 0u);
}

__global__ void philox_words_kernel(PhiloxKey key, uint32_t tag, int64_t first, int64_t n,
                                    uint32_t *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = element_block(key, tag, first + i);
    reinterpret_cast<uint4 *>(out)[i] = r;
  }
}

__global__ void uniform_kernel(PhiloxKey key, uint32_t tag, int64_t first, int64_t n,
                               double *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 r = element_block(key, tag, first + i);
    out[i] = u53(r.x, r.y);
  }
}

__global__ void normal_kernel(PhiloxKey key, uint32_t tag, int64_t first, int64_t n, double loc0,
                              double scale0, const double *__restrict__ loc,
                              const double *__restrict__ scale, double *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double z = normal_f64(element_block(key, tag, first + i));
    const double m = loc ? loc[i] : loc0, s = scale ? scale[i] : scale0;
    out[i] = m + s * z;                     // numpy legacy: loc + scale * gauss
  }
}

// searchsorted(cdf, u, side='right'): first index with cdf[idx] > u
__device__ __forceinline__ int upper_bound(const double *__restrict__ cdf, int n, double u) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void density_index_kernel(const double *__restrict__ cdf, int P,
                                     const double *__restrict__ u, int64_t n,
                                     const int32_t *__restrict__ pos_pt,
                                     const int32_t *__restrict__ pos_branch,
                                     int32_t *__restrict__ row_of_cell, int64_t *__restrict__ pt,
                                     int32_t *__restrict__ branch) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int idx = upper_bound(cdf, P, u[i]);
    idx = min(idx, P - 1);                  // cdf[-1] == 1.0 > u always; guard anyway
    row_of_cell[i] = idx;
    if (pt) pt[i] = pos_pt[idx];
    if (branch) branch[i] = pos_branch[idx];
  }
}

__global__ void times_kernel(const double *__restrict__ z, int64_t n, int max_time,
                             int64_t *__restrict__ pt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    // simulation.py:410-412: astype(int) truncates toward zero, then clip
    double v = trunc(z[i]);
    v = fmin(fmax(v, 0.0), (double)(max_time - 1));
    pt[i] = (int64_t)v;
  }
}

constexpr int MAX_CAND = 64;   // live branches per timezone handled on device

__global__ void pick_branch_kernel(const int64_t *__restrict__ pt, const double *__restrict__ u,
                                   int64_t n, int nz, const int32_t *__restrict__ zone_lo,
                                   const int32_t *__restrict__ zone_hi,
                                   const int32_t *__restrict__ cand_off,
                                   const int32_t *__restrict__ cand_branch,
                                   const int32_t *__restrict__ branch_start,
                                   const int32_t *__restrict__ row_base,
                                   const int32_t *__restrict__ T, const double *__restrict__ density,
                                   int32_t *__restrict__ branch, int32_t *__restrict__ row_of_cell,
                                   uint32_t *flags) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = pt[i];
    int z = -1;
    for (int q = 0; q < nz; ++q)            // sim_utils.py:388-391: first zone containing t
      if (t >= zone_lo[q] && t <= zone_hi[q]) { z = q; break; }
    if (z < 0) { atomicOr(flags, PST_FLAG_NOZONE); branch[i] = -1; row_of_cell[i] = -1; continue; }
    const int c0 = cand_off[z], nc = cand_off[z + 1] - c0;
    const int where = (int)(t - zone_lo[z]);            // sim_utils.py:393 (sic, Q5)
    double w[MAX_CAND];
    bool bad = false;
    for (int c = 0; c < nc; ++c) {
      const int b = cand_branch[c0 + c];
      if (where >= T[b]) { bad = true; w[c] = 0.0; continue; }
      w[c] = density[row_base[b] + where];              // :396
    }
    // densities.sum() in numpy's order: sequential below 8 terms, else 8 running
    // partial sums combined pairwise, remainder added last (nc <= 64 < 128)
    double tot = 0.0;
    if (nc < 8) {
      for (int c = 0; c < nc; ++c) tot += w[c];
    } else {
      double r[8];
      for (int q = 0; q < 8; ++q) r[q] = w[q];
      int c = 8;
      for (; c < nc - (nc % 8); c += 8)
        for (int q = 0; q < 8; ++q) r[q] += w[c + q];
      tot = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
      for (; c < nc; ++c) tot += w[c];
    }
    if (bad || nc == 0) { atomicOr(flags, PST_FLAG_NOZONE); branch[i] = -1; row_of_cell[i] = -1; continue; }
    // :397-399 -> legacy choice: cdf = cumsum(p); cdf /= cdf[-1]; searchsorted right
    double acc = 0.0;
    for (int c = 0; c < nc; ++c) { acc += w[c] / tot; w[c] = acc; }
    const double last = w[nc - 1];
    int pick = nc - 1;
    for (int c = 0; c < nc; ++c) if (w[c] / last > u[i]) { pick = c; break; }
    const int b = cand_branch[c0 + pick];
    branch[i] = b;
    const int64_t rel = t - branch_start[b];
    if (rel < 0 || rel >= T[b]) { atomicOr(flags, PST_FLAG_ROW); row_of_cell[i] = -1; }
    else row_of_cell[i] = row_base[b] + (int)rel;
  }
}

__global__ void rows_from_branch_kernel(const int64_t *__restrict__ pt,
                                        const int32_t *__restrict__ branch, int64_t n, int B,
                                        const int32_t *__restrict__ branch_start,
                                        const int32_t *__restrict__ row_base,
                                        const int32_t *__restrict__ T,
                                        int32_t *__restrict__ row_of_cell, uint32_t *flags) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int b = branch[i];
    int row = -1;
    if (b >= 0 && b < B) {
      const int64_t rel = pt[i] - branch_start[b];      // simulation.py:634-635
      if (rel >= 0 && rel < T[b]) row = row_base[b] + (int)rel;
    }
    if (row < 0) atomicOr(flags, PST_FLAG_ROW);
    row_of_cell[i] = row;
  }
}

__global__ void whole_tree_kernel(const int32_t *__restrict__ cover_pt,
                                  const int32_t *__restrict__ cover_branch,
                                  const int32_t *__restrict__ cover_row, int64_t first, int64_t n,
                                  int64_t n_factor, int64_t *__restrict__ pt,
                                  int32_t *__restrict__ branch, int32_t *__restrict__ row_of_cell) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = (first + i) / n_factor;           // np.repeat, simulation.py:512-513
    pt[i] = cover_pt[e];
    branch[i] = cover_branch[e];
    row_of_cell[i] = cover_row[e];
  }
}

__global__ void scalings_kernel(const double *__restrict__ z, int64_t n, double *__restrict__ out64,
                                float *__restrict__ out32) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double s = z ? exp(z[i]) : 1.0;               // sim_utils.py:494-497
    if (out64) out64[i] = s;
    if (out32) out32[i] = (float)s;
  }
}

// sim_utils.py:463-469 with the redraws as counter-based attempts: attempt a of gene g reads the
// normal at element a*G + g, so the result does not depend on the launch shape
__global__ void base_gene_exp_kernel(PhiloxKey key, uint32_t tag, const double *__restrict__ cap, int64_t G,
                                     double abs_max, double gene_mean, double gene_std, int max_tries,
                                     double *__restrict__ out_base, int32_t *__restrict__ out_tries,
                                     uint32_t *__restrict__ flags) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < G;
       g += (int64_t)gridDim.x * blockDim.x) {
    const double c = cap[g];
    double value = 0.0;
    int a = 0;
    for (; a < max_tries; ++a) {
      value = exp(gene_mean + gene_std * normal_f64(element_block(key, tag, (int64_t)a * G + g)));
      if (!(value * c > abs_max)) break;
    }
    if (a == max_tries) atomicOr(flags, (uint32_t)PST_FLAG_DOMAIN);
    out_base[g] = value;
    if (out_tries) out_tries[g] = a + 1;
  }
}

__global__ void nb_params_kernel(const double *__restrict__ alpha, const double *__restrict__ beta,
                                 const double *__restrict__ mu, int64_t n_cells, int64_t G,
                                 double *__restrict__ out_p, double *__restrict__ out_r) {
  const int64_t n = n_cells * G;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = i % G;
    const double m = mu[i], a = alpha[g], b = beta[g];
    const double s2 = a * m * m + b * m;                // count_model.py:156
    double p = (s2 - m) / s2, r = (m * m) / (s2 - m);   // :157-158
    if (s2 <= 0.0) { p = 0.0; r = 0.0; }                // :159-160
    out_p[i] = p;
    out_r[i] = r;
  }
}

static inline unsigned grid_for(int64_t n) {
  const int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)num_sm() * 8;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace pst

using namespace pst;

extern "C" int pst_abi_version(void) { return PST_ABI_VERSION; }
extern "C" const char *pst_last_error(void) { return g_err; }
extern "C" uint64_t pst_launch_count(void) { return g_launches.load(); }

extern "C" int pst_philox_words(uint64_t seed, uint32_t tag, int64_t first, int64_t n, uint32_t *out,
                                void *stream) {
  const char *fn = "pst_philox_words";
  PST_REQUIRE(n >= 0 && first >= 0, fn, "negative size");
  if (n == 0) return 0;
  PST_REQUIRE(out, fn, "null pointer");
  philox_words_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(PhiloxKey(seed), tag, first, n, out);
  return check_launch(fn);
}

extern "C" int pst_uniform_f64(uint64_t seed, uint32_t tag, int64_t first, int64_t n, double *out,
                               void *stream) {
  const char *fn = "pst_uniform_f64";
  PST_REQUIRE(n >= 0 && first >= 0, fn, "negative size");
  if (n == 0) return 0;
  PST_REQUIRE(out, fn, "null pointer");
  uniform_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(PhiloxKey(seed), tag, first, n, out);
  return check_launch(fn);
}

extern "C" int pst_normal_f64(uint64_t seed, uint32_t tag, int64_t first, int64_t n, double loc0,
                              double scale0, const double *loc, const double *scale, double *out,
                              void *stream) {
  const char *fn = "pst_normal_f64";
  PST_REQUIRE(n >= 0 && first >= 0, fn, "negative size");
  if (n == 0) return 0;
  PST_REQUIRE(out, fn, "null pointer");
  normal_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(PhiloxKey(seed), tag, first, n, loc0, scale0, loc, scale, out);
  return check_launch(fn);
}

extern "C" int pst_density_index(const double *cdf, int32_t P, const double *u, int64_t n,
                                 const int32_t *pos_pt, const int32_t *pos_branch,
                                 int32_t *row_of_cell, int64_t *pt, int32_t *branch, void *stream) {
  const char *fn = "pst_density_index";
  PST_REQUIRE(n >= 0, fn, "negative size");
  if (n == 0) return 0;
  PST_REQUIRE(P > 0, fn, "empty tree (P == 0)");
  PST_REQUIRE(cdf && u && row_of_cell, fn, "null pointer");
  PST_REQUIRE((!pt || pos_pt) && (!branch || pos_branch), fn, "position tables missing");
  density_index_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(cdf, P, u, n, pos_pt, pos_branch,
                                                                      row_of_cell, pt, branch);
  return check_launch(fn);
}

extern "C" int pst_times_from_normals(const double *z, int64_t n, int32_t max_time, int64_t *pt,
                                      void *stream) {
  const char *fn = "pst_times_from_normals";
  PST_REQUIRE(n >= 0, fn, "negative size");
  if (n == 0) return 0;
  PST_REQUIRE(max_time > 0, fn, "max_time must be positive");
  PST_REQUIRE(z && pt, fn, "null pointer");
  times_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(z, n, max_time, pt);
  return check_launch(fn);
}

extern "C" int pst_pick_branch(const int64_t *pt, const double *u, int64_t n, int32_t nz,
                               const int32_t *zone_lo, const int32_t *zone_hi,
                               const int32_t *cand_off, const int32_t *cand_branch,
                               int32_t max_cand, const int32_t *branch_start,
                               const int32_t *row_base, const int32_t *T, const double *density,
                               int32_t *branch, int32_t *row_of_cell, uint32_t *flags,
                               void *stream) {
  const char *fn = "pst_pick_branch";
  PST_REQUIRE(n >= 0 && nz >= 0, fn, "negative size");
  if (n == 0) return 0;
  PST_REQUIRE(max_cand <= MAX_CAND, fn, "more than 64 live branches in one timezone");
  PST_REQUIRE(pt && u && zone_lo && zone_hi && cand_off && cand_branch && branch_start && row_base &&
                  T && density && branch && row_of_cell && flags, fn, "null pointer");
  pick_branch_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(
      pt, u, n, nz, zone_lo, zone_hi, cand_off, cand_branch, branch_start, row_base, T, density, branch,
      row_of_cell, flags);
  return check_launch(fn);
}

extern "C" int pst_rows_from_branch(const int64_t *pt, const int32_t *branch, int64_t n, int32_t B,
                                    const int32_t *branch_start, const int32_t *row_base,
                                    const int32_t *T, int32_t *row_of_cell, uint32_t *flags,
                                    void *stream) {
  const char *fn = "pst_rows_from_branch";
  PST_REQUIRE(n >= 0 && B >= 0, fn, "negative size");
  if (n == 0) return 0;
  PST_REQUIRE(pt && branch && branch_start && row_base && T && row_of_cell && flags, fn, "null pointer");
  rows_from_branch_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(pt, branch, n, B, branch_start,
                                                                         row_base, T, row_of_cell, flags);
  return check_launch(fn);
}

extern "C" int pst_whole_tree_index(const int32_t *cover_pt, const int32_t *cover_branch,
                                    const int32_t *cover_row, int64_t n_cover, int64_t n_factor,
                                    int64_t first, int64_t n, int64_t *pt, int32_t *branch,
                                    int32_t *row_of_cell, void *stream) {
  const char *fn = "pst_whole_tree_index";
  PST_REQUIRE(n >= 0 && first >= 0 && n_cover >= 0, fn, "negative size");
  if (n == 0) return 0;
  PST_REQUIRE(n_factor > 0, fn, "n_factor must be positive");
  PST_REQUIRE(first + n <= n_cover * n_factor, fn, "cell range exceeds n_factor * positions");
  PST_REQUIRE(cover_pt && cover_branch && cover_row && pt && branch && row_of_cell, fn, "null pointer");
  whole_tree_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(cover_pt, cover_branch, cover_row, first,
                                                                   n, n_factor, pt, branch, row_of_cell);
  return check_launch(fn);
}

extern "C" int pst_scalings(const double *z, int64_t n, double *out64, float *out32, void *stream) {
  const char *fn = "pst_scalings";
  PST_REQUIRE(n >= 0, fn, "negative size");
  if (n == 0) return 0;
  PST_REQUIRE(out64 || out32, fn, "no output");
  scalings_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(z, n, out64, out32);
  return check_launch(fn);
}

extern "C" int pst_base_gene_exp(uint64_t seed, uint32_t tag, const double *cap, int64_t G, double abs_max,
                                 double gene_mean, double gene_std, int32_t max_tries, double *out_base,
                                 int32_t *out_tries, uint32_t *flags, void *stream) {
  const char *fn = "pst_base_gene_exp";
  PST_REQUIRE(G >= 0, fn, "negative size");
  if (G == 0) return 0;
  PST_REQUIRE(cap && out_base && flags, fn, "null pointer");
  PST_REQUIRE(max_tries > 0, fn, "max_tries must be positive");
  base_gene_exp_kernel<<<grid_for(G), 256, 0, (cudaStream_t)stream>>>(PhiloxKey(seed), tag, cap, G, abs_max,
                                                                      gene_mean, gene_std, max_tries, out_base,
                                                                      out_tries, flags);
  return check_launch(fn);
}

extern "C" int pst_nb_params(const double *alpha, const double *beta, const double *mu,
                             int64_t n_cells, int64_t G, double *out_p, double *out_r, void *stream) {
  const char *fn = "pst_nb_params";
  PST_REQUIRE(n_cells >= 0 && G >= 0, fn, "negative size");
  if (n_cells == 0 || G == 0) return 0;
  PST_REQUIRE(alpha && beta && mu && out_p && out_r, fn, "null pointer");
  nb_params_kernel<<<grid_for(n_cells * G), 256, 0, (cudaStream_t)stream>>>(alpha, beta, mu, n_cells, G,
                                                                            out_p, out_r);
  return check_launch(fn);
}
