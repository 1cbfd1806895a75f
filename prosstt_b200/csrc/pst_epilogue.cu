// Epilogues over the count matrix that every reference notebook runs on the host right after
// sampling: library-size normalisation and log transform
// (compare_axolotl.ipynb cell 14: `(X.T / scalings).T`, minimal_example.ipynb cell 6:
// `np.log(X + 1)`) and compaction of the (N, G) matrix into CSR for the on-disk formats
// single-cell tools read (tree_utils.py:59-173 writes dense text only).
//
// Both are pure streaming passes (HBM-bound): one 128-bit load per gene quad, rows walked by
// warps so that every warp touches 512 contiguous bytes.
#include <algorithm>
#include "pst_common.cuh"

namespace pst {
namespace {

// ---------------------------------------------------------------------------------------------
// out[i][g] = f(X[i][g], scaling[i]);  mode 0: X/s   1: log(X/s + 1)   2: log(X + 1)
// ---------------------------------------------------------------------------------------------
// grid.x walks the gene quads of a row, grid.y strides over rows: no 64-bit division per item
template <int MODE, bool VEC>
__global__ void __launch_bounds__(256)
transform_kernel(const int32_t *__restrict__ X, int64_t n, int64_t G, int64_t ldx,
                 const float *__restrict__ scaling, float *__restrict__ out, int64_t ldo) {
  const int64_t Q = (G + 3) / 4;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g0 = q * 4;
    for (int64_t row = blockIdx.y; row < n; row += gridDim.y) {
      // one correctly rounded reciprocal per item, then multiplies: within 1.5 ulp of x / s
      const float inv = (MODE == PST_TRANSFORM_LOG1P) ? 1.0f : 1.0f / __ldg(scaling + row);
      const int32_t *src = X + row * ldx + g0;
      float *dst = out + row * ldo + g0;
      if (VEC) {
        const int4 v = __ldcs(reinterpret_cast<const int4 *>(src));
        float4 o = make_float4((float)v.x * inv, (float)v.y * inv, (float)v.z * inv, (float)v.w * inv);
        if (MODE != PST_TRANSFORM_NORMALIZE)
          o = make_float4(log1pf(o.x), log1pf(o.y), log1pf(o.z), log1pf(o.w));
        __stcs(reinterpret_cast<float4 *>(dst), o);
      } else {
        for (int j = 0; j < 4 && g0 + j < G; ++j) {
          const float o = (float)src[j] * inv;
          dst[j] = (MODE != PST_TRANSFORM_NORMALIZE) ? log1pf(o) : o;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// CSR fill: one warp per row, 128 genes per step; a lane owns a gene quad, so column order inside
// the row is preserved by an exclusive warp prefix of the lanes' nonzero counts
// ---------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256)
csr_fill_kernel(const int32_t *__restrict__ X, int64_t n, int64_t G, int64_t ldx,
                const int64_t *__restrict__ indptr, int32_t *__restrict__ indices,
                int32_t *__restrict__ data, uint32_t *__restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t Q = (G + 3) / 4;
  for (int64_t row = warp0; row < n; row += n_warps) {
    int64_t pos = indptr[row];
    const int64_t end = indptr[row + 1];
    const int32_t *src = X + row * ldx;
#pragma unroll 4
    for (int64_t q0 = 0; q0 < Q; q0 += 32) {
      const int64_t q = q0 + lane;
      int4 v = make_int4(0, 0, 0, 0);
      if (q < Q) {
        if (VEC) {
          v = __ldcs(reinterpret_cast<const int4 *>(src + q * 4));
        } else {
          const int64_t g = q * 4;
          v.x = src[g];
          if (g + 1 < G) v.y = src[g + 1];
          if (g + 2 < G) v.z = src[g + 2];
          if (g + 3 < G) v.w = src[g + 3];
        }
      }
      // exclusive prefix of the lanes' nonzero counts (0..4) from three ballots of the count's bits
      const int c = (v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0);
      const unsigned b0 = __ballot_sync(0xffffffffu, c & 1), b1 = __ballot_sync(0xffffffffu, c & 2),
                     b2 = __ballot_sync(0xffffffffu, c & 4);
      const int excl = __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
      const int total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
      int64_t w = pos + excl;
      if (pos + total <= end) {                   // indptr is consistent with the matrix
        const int32_t g = (int32_t)(q * 4);
        if (v.x != 0) { indices[w] = g;     data[w] = v.x; ++w; }
        if (v.y != 0) { indices[w] = g + 1; data[w] = v.y; ++w; }
        if (v.z != 0) { indices[w] = g + 2; data[w] = v.z; ++w; }
        if (v.w != 0) { indices[w] = g + 3; data[w] = v.w; ++w; }
      }
      pos += total;
    }
    if (lane == 0 && pos != end) atomicOr(flags, (uint32_t)PST_FLAG_ROW);   // indptr does not match X
  }
}

// ---------------------------------------------------------------------------------------------
// narrow transfer formats: out = min(X, SAT) as uint8 (SAT = 255) or uint16 (SAT = 65535); every
// element that reads SAT also has an entry (flat index, exact value) in the overflow list, so the
// int32 matrix can be rebuilt exactly
// ---------------------------------------------------------------------------------------------
template <typename OUT, bool VEC>
__global__ void __launch_bounds__(256)
narrow_kernel(const int32_t *__restrict__ X, int64_t n, int64_t G, int64_t ldx, OUT *__restrict__ out, int64_t ldo,
              int64_t row0, int64_t *__restrict__ ovf_index, int32_t *__restrict__ ovf_value, int64_t ovf_cap,
              unsigned long long *__restrict__ ovf_count) {
  constexpr int SAT = (sizeof(OUT) == 1) ? 255 : 65535;
  const int64_t Q = (G + 3) / 4;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < Q; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g0 = q * 4;
    for (int64_t row = blockIdx.y; row < n; row += gridDim.y) {
      const int32_t *src = X + row * ldx + g0;
      OUT *dst = out + row * ldo + g0;
      int v[4] = {0, 0, 0, 0};
      if (VEC) {
        const int4 w = __ldcs(reinterpret_cast<const int4 *>(src));
        v[0] = w.x; v[1] = w.y; v[2] = w.z; v[3] = w.w;
      } else {
        for (int j = 0; j < 4 && g0 + j < G; ++j) v[j] = src[j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (v[j] >= SAT && g0 + j < G) {                              // rare: one atomic per listed count
          const unsigned long long slot = atomicAdd(ovf_count, 1ull);
          if ((int64_t)slot < ovf_cap) { ovf_index[slot] = (row0 + row) * G + g0 + j; ovf_value[slot] = v[j]; }
          v[j] = SAT;
        }
      }
      if (VEC) {
        if (sizeof(OUT) == 1) {
          __stcs(reinterpret_cast<unsigned *>(dst),
                 (unsigned)v[0] | ((unsigned)v[1] << 8) | ((unsigned)v[2] << 16) | ((unsigned)v[3] << 24));
        } else {
          __stcs(reinterpret_cast<uint2 *>(dst), make_uint2((unsigned)v[0] | ((unsigned)v[1] << 16),
                                                            (unsigned)v[2] | ((unsigned)v[3] << 16)));
        }
      } else {
        for (int j = 0; j < 4 && g0 + j < G; ++j) dst[j] = (OUT)v[j];
      }
    }
  }
}

// Pure-store kernel: the write-only HBM ceiling the count write is measured against.  Same
// shape of traffic as the sampler's output: 128-bit stores, 512 contiguous bytes per warp instruction.
__global__ void __launch_bounds__(256) store_fill_kernel(int4 *__restrict__ out, int64_t n16, int32_t value) {
  const int4 v = make_int4(value, value, value, value);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) out[i] = v;
}

inline unsigned stream_grid(int64_t threads) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>((threads + 255) / 256, (int64_t)num_sm() * 16));
}

}  // namespace
}  // namespace pst

using namespace pst;

extern "C" int pst_transform_counts(const int32_t *X, int64_t n, int64_t G, int64_t ldx, const float *scaling,
                                    int32_t mode, float *out, int64_t ldo, void *stream) {
  const char *fn = "pst_transform_counts";
  PST_REQUIRE(n >= 0 && G >= 0 && ldx >= G && ldo >= G, fn, "need n, G >= 0, ldx >= G and ldo >= G");
  PST_REQUIRE(mode == PST_TRANSFORM_NORMALIZE || mode == PST_TRANSFORM_NORMALIZE_LOG1P ||
              mode == PST_TRANSFORM_LOG1P, fn, "unknown mode");
  if (n == 0 || G == 0) return 0;
  PST_REQUIRE(X && out, fn, "null pointer");
  PST_REQUIRE(scaling || mode == PST_TRANSFORM_LOG1P, fn, "scaling is required for the normalising modes");
  const bool vec = (G % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && ((uintptr_t)X % 16 == 0) &&
                   ((uintptr_t)out % 16 == 0);
  const int64_t qblocks = std::min<int64_t>((((G + 3) / 4) + 255) / 256, 64);
  const int64_t yblocks = std::max<int64_t>(1, std::min<int64_t>(n, ((int64_t)num_sm() * 16 + qblocks - 1) / qblocks));
  const dim3 grid((unsigned)qblocks, (unsigned)yblocks);
  cudaStream_t st = (cudaStream_t)stream;
#define PST_LAUNCH_TRANSFORM(M)                                                                  \
  do {                                                                                           \
    if (vec) transform_kernel<M, true><<<grid, 256, 0, st>>>(X, n, G, ldx, scaling, out, ldo);   \
    else transform_kernel<M, false><<<grid, 256, 0, st>>>(X, n, G, ldx, scaling, out, ldo);      \
  } while (0)
  if (mode == PST_TRANSFORM_NORMALIZE) PST_LAUNCH_TRANSFORM(PST_TRANSFORM_NORMALIZE);
  else if (mode == PST_TRANSFORM_NORMALIZE_LOG1P) PST_LAUNCH_TRANSFORM(PST_TRANSFORM_NORMALIZE_LOG1P);
  else PST_LAUNCH_TRANSFORM(PST_TRANSFORM_LOG1P);
#undef PST_LAUNCH_TRANSFORM
  return check_launch(fn);
}

extern "C" int pst_store_fill(int32_t *out, int64_t n, int32_t value, void *stream) {
  const char *fn = "pst_store_fill";
  PST_REQUIRE(n >= 0 && n % 4 == 0, fn, "n must be a non-negative multiple of 4");
  if (n == 0) return 0;
  PST_REQUIRE(out && (uintptr_t)out % 16 == 0, fn, "out must be a 16-byte aligned device pointer");
  store_fill_kernel<<<(unsigned)((int64_t)num_sm() * 8), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<int4 *>(out), n / 4, value);
  return check_launch(fn);
}

extern "C" int pst_csr_fill(const int32_t *X, int64_t n, int64_t G, int64_t ldx, const int64_t *indptr,
                            int32_t *indices, int32_t *data, uint32_t *flags, void *stream) {
  const char *fn = "pst_csr_fill";
  PST_REQUIRE(n >= 0 && G >= 0 && ldx >= G, fn, "need n, G >= 0 and ldx >= G");
  PST_REQUIRE(G < ((int64_t)1 << 31), fn, "column index does not fit int32");
  if (n == 0 || G == 0) return 0;
  PST_REQUIRE(X && indptr && indices && data && flags, fn, "null pointer");
  const bool vec = (G % 4 == 0) && (ldx % 4 == 0) && ((uintptr_t)X % 16 == 0);
  const unsigned grid = stream_grid(n * 32);
  if (vec) csr_fill_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(X, n, G, ldx, indptr, indices, data, flags);
  else csr_fill_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(X, n, G, ldx, indptr, indices, data, flags);
  return check_launch(fn);
}

extern "C" int pst_narrow_counts(const int32_t *X, int64_t n, int64_t G, int64_t ldx, void *out, int64_t ldo,
                                 int32_t out_bits, int64_t row0, int64_t *ovf_index, int32_t *ovf_value,
                                 int64_t ovf_cap, uint64_t *ovf_count, void *stream) {
  const char *fn = "pst_narrow_counts";
  PST_REQUIRE(n >= 0 && G >= 0 && ldx >= G && ldo >= G && row0 >= 0 && ovf_cap >= 0, fn,
              "need n, G, row0, ovf_cap >= 0, ldx >= G and ldo >= G");
  PST_REQUIRE(out_bits == 8 || out_bits == 16, fn, "out_bits must be 8 or 16");
  if (n == 0 || G == 0) return 0;
  PST_REQUIRE(X && out && ovf_count && (ovf_cap == 0 || (ovf_index && ovf_value)), fn, "null pointer");
  const int64_t obytes = out_bits / 8;
  const bool vec = (G % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && ((uintptr_t)X % 16 == 0) &&
                   ((uintptr_t)out % (4 * obytes) == 0);
  const int64_t qblocks = std::min<int64_t>((((G + 3) / 4) + 255) / 256, 64);
  const int64_t yblocks = std::max<int64_t>(1, std::min<int64_t>(n, ((int64_t)num_sm() * 16 + qblocks - 1) / qblocks));
  const dim3 grid((unsigned)qblocks, (unsigned)yblocks);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long *cnt = (unsigned long long *)ovf_count;
#define PST_LAUNCH_NARROW(T, V) \
  narrow_kernel<T, V><<<grid, 256, 0, st>>>(X, n, G, ldx, (T *)out, ldo, row0, ovf_index, ovf_value, ovf_cap, cnt)
  if (out_bits == 8) { if (vec) PST_LAUNCH_NARROW(uint8_t, true); else PST_LAUNCH_NARROW(uint8_t, false); }
  else { if (vec) PST_LAUNCH_NARROW(uint16_t, true); else PST_LAUNCH_NARROW(uint16_t, false); }
#undef PST_LAUNCH_NARROW
  return check_launch(fn);
}
