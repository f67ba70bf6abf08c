// Eval-epoch metrics on the device: the per-batch top-k / small-k samples of every unit and
// their merge into the running top-k (model_pipeline.py:335-360, utils.py:1445-1481), the spatial means they are taken
// over (utils.py:1996-2010) and the activation histograms (utils.py:1934-1963).  All integer / index work: the indices
// these kernels return are the ones torch.topk returns for the same values (ties: lower row first), bit for bit.
#include "svb_common.cuh"

using namespace svb;

namespace {

// ------------------------------------------------------------------------------------------------ spatial mean
// Token-major t [n_images * hw, F] -> out [n_images, F]: mean over the hw rows of every image.  grid (F tiles, images),
// 256 threads: lane owns kN consecutive columns (16-byte loads), warp w owns rows w, w+8, ...; the eight warp partials
// are added in a fixed order.
template <typename T>
__global__ void __launch_bounds__(256)
spatial_mean_tokens_kernel(const T* __restrict__ t, float* __restrict__ out, int hw, int F) {
  constexpr int V = Vec16<T>::kN;
  __shared__ float s[8][32 * V + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int f0 = (blockIdx.x * 32 + lane) * V;
  const long long row0 = static_cast<long long>(blockIdx.y) * hw;
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  if (f0 < F) {
    for (int r = w; r < hw; r += 8) {
      float v[8];
      Vec16<T>::load(t + (row0 + r) * F + f0, v);
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += v[k];
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) s[w][lane * V + k] = acc[k];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * V; c += 256) {
    const int f = blockIdx.x * 32 * V + c;
    if (f < F) {
      float sum = 0.f;
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) sum += s[ww][c];
      out[static_cast<size_t>(blockIdx.y) * F + f] = sum / static_cast<float>(hw);
    }
  }
}
// [n_images, F, hw] (NCHW) -> out [n_images, F]: one warp per (image, unit) row of hw contiguous values.
template <typename T>
__global__ void __launch_bounds__(256)
spatial_mean_nchw_kernel(const T* __restrict__ t, float* __restrict__ out, long long rows, int hw) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int p = lane; p < hw; p += 32) acc += to_f32(t[row * hw + p]);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = acc / static_cast<float>(hw);
}

// Sum over images: t [n_images, R, F] (token-major rows of R positions per image, or R = 1 with F = C*HW for an NCHW
// tensor) -> out [R, F] = sum_b t[b].  One thread per 16 bytes of a row, images in order (deterministic); the running
// sums of compute_ie.py:146-207 (encoder output, SAE error, layer output per position) are these.
template <typename T>
__global__ void __launch_bounds__(256)
image_sum_kernel(const T* __restrict__ t, float* __restrict__ out, long long n_images, long long per_image) {
  constexpr int V = Vec16<T>::kN;
  const long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * V;
  if (i >= per_image) return;
  float acc[8];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  for (long long b = 0; b < n_images; ++b) {
    float v[8];
    Vec16<T>::load(t + b * per_image + i, v);
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] += v[k];
  }
#pragma unroll
  for (int k = 0; k < V; ++k) out[i + k] = acc[k];
}

// ------------------------------------------------------------------------------------------------ top-k over rows, per column
// Candidates of column f: rows 0..n0-1 of source 0 followed by rows 0..n1-1 of source 1 (the running top-k and the
// batch's, utils.py:1463-1467; n1 = 0 for a plain per-batch top-k).  A 64-bit key (value mapped to an order-preserving
// unsigned | candidate position) is bitonic-sorted in shared memory, so equal values keep the lower position first and
// the result does not depend on the launch configuration.  NaN sorts as the largest value, like torch.topk.
__device__ __forceinline__ uint32_t orderable(float v) {
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void __launch_bounds__(256)
topk_columns_kernel(const float* __restrict__ v0, const int64_t* __restrict__ i0, const int64_t* __restrict__ f0, int n0,
                    const float* __restrict__ v1, const int64_t* __restrict__ i1, const int64_t* __restrict__ f1, int n1,
                    int F, int k, int largest, int P, float* __restrict__ out_v, int64_t* __restrict__ out_i,
                    int64_t* __restrict__ out_f) {
  extern __shared__ unsigned long long keys[];
  const int col = blockIdx.x;
  const int n = n0 + n1;
  for (int c = threadIdx.x; c < P; c += blockDim.x) {
    unsigned long long key = ~0ull;                      // padding sorts last
    if (c < n) {
      const float v = c < n0 ? v0[static_cast<size_t>(c) * F + col] : v1[static_cast<size_t>(c - n0) * F + col];
      uint32_t o = orderable(v);
      if (largest) o = ~o;
      key = (static_cast<unsigned long long>(o) << 32) | static_cast<uint32_t>(c);
    }
    keys[c] = key;
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int c = threadIdx.x; c < P / 2; c += blockDim.x) {
        const int lo = 2 * c - (c & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const int c = static_cast<int>(keys[j] & 0xffffffffu);
    const bool s0 = c < n0;
    const size_t src = static_cast<size_t>(s0 ? c : c - n0) * F + col;
    const size_t dst = static_cast<size_t>(j) * F + col;
    out_v[dst] = s0 ? v0[src] : v1[src];
    if (out_i) out_i[dst] = s0 ? (i0 ? i0[src] : c) : (i1 ? i1[src] : c - n0);
    if (out_f) out_f[dst] = s0 ? (f0 ? f0[src] : 0) : (f1 ? f1[src] : 0);
  }
}

// ------------------------------------------------------------------------------------------------ histograms
// hist[bin, u] += #{rows r : vals[r, unit_idx[u]] falls into bin} with torch.histc's CUDA binning (values outside
// [min, max] are ignored, max itself belongs to the last bin; min == max means "use the data's own minimum and
// maximum", and only if those coincide too the range is widened by one on both sides).
__global__ void __launch_bounds__(256)
histogram_columns_kernel(const float* __restrict__ vals, long long rows, int F, const int64_t* __restrict__ unit_idx,
                         const float* __restrict__ mins, const float* __restrict__ maxs, int bins,
                         float* __restrict__ hist, int U) {
  extern __shared__ int counts[];
  const int u = blockIdx.x;
  for (int b = threadIdx.x; b < bins; b += blockDim.x) counts[b] = 0;
  __syncthreads();
  const long long col = unit_idx ? unit_idx[u] : u;
  float lo = mins[u], hi = maxs[u];
  if (lo == hi) {   // block-uniform branch: the column's own range
    __shared__ float smin[8], smax[8];
    float a = INFINITY, b = -INFINITY;
    for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
      const float x = vals[r * F + col];
      a = fminf(a, x); b = fmaxf(b, x);
    }
    for (int o = 16; o > 0; o >>= 1) {
      a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = a; smax[threadIdx.x >> 5] = b; }
    __syncthreads();
    lo = smin[0]; hi = smax[0];
    for (int w = 1; w < 8; ++w) { lo = fminf(lo, smin[w]); hi = fmaxf(hi, smax[w]); }
    if (lo == hi) { lo -= 1.f; hi += 1.f; }
  }
  for (long long r = threadIdx.x; r < rows; r += blockDim.x) {
    const float x = vals[r * F + col];
    if (x >= lo && x <= hi) {
      int bin = static_cast<int>((x - lo) * static_cast<float>(bins) / (hi - lo));
      if (bin == bins) bin -= 1;
      atomicAdd(&counts[bin], 1);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < bins; b += blockDim.x) hist[static_cast<size_t>(b) * U + u] += static_cast<float>(counts[b]);
}

}  // namespace

extern "C" int svb_spatial_mean(svb_handle* h, void* stream, const void* t, int32_t dtype, int32_t layout,
                                int64_t n_images, int32_t hw, int32_t F, float* out) {
  if (!h || !t || !out) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (n_images <= 0 || hw <= 0 || F <= 0) return fail(SVB_ERR_BAD_ARG, "empty tensor");
  if (dtype != SVB_F32 && dtype != SVB_BF16) return fail(SVB_ERR_BAD_ARG, "bad dtype %d", dtype);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (layout == SVB_NCHW) {
    const long long rows = n_images * static_cast<long long>(F);
    if (dtype == SVB_F32)
      (spatial_mean_nchw_kernel<float><<<cdiv(rows, 8), 256, 0, st>>>(static_cast<const float*>(t), out, rows, hw), svb::count_launch());
    else
      (spatial_mean_nchw_kernel<bf16><<<cdiv(rows, 8), 256, 0, st>>>(static_cast<const bf16*>(t), out, rows, hw), svb::count_launch());
  } else {
    const int V = dtype == SVB_F32 ? 4 : 8;
    if (F % V || (reinterpret_cast<uintptr_t>(t) & 15))
      return fail(SVB_ERR_UNSUPPORTED, "token-major input needs F %% %d == 0 and a 16-byte aligned pointer", V);
    if (n_images > 65535) return fail(SVB_ERR_UNSUPPORTED, "more than 65535 images per call");
    const dim3 grid(cdiv(F, 32 * V), static_cast<unsigned>(n_images));
    if (dtype == SVB_F32)
      (spatial_mean_tokens_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(t), out, hw, F), svb::count_launch());
    else
      (spatial_mean_tokens_kernel<bf16><<<grid, 256, 0, st>>>(static_cast<const bf16*>(t), out, hw, F), svb::count_launch());
  }
  SVB_LAUNCH_CHECK("spatial_mean");
  return 0;
}

extern "C" int svb_image_sum(svb_handle* h, void* stream, const void* t, int32_t dtype, int64_t n_images,
                             int64_t per_image, float* out) {
  if (!h || !t || !out) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (n_images <= 0 || per_image <= 0) return fail(SVB_ERR_BAD_ARG, "empty tensor");
  const int V = dtype == SVB_F32 ? 4 : 8;
  if (dtype != SVB_F32 && dtype != SVB_BF16) return fail(SVB_ERR_BAD_ARG, "bad dtype %d", dtype);
  if (per_image % V || (reinterpret_cast<uintptr_t>(t) & 15))
    return fail(SVB_ERR_UNSUPPORTED, "svb_image_sum needs %d | elements per image and a 16-byte aligned pointer", V);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((per_image / V + 255) / 256);
  if (dtype == SVB_F32)
    (image_sum_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(t), out, n_images, per_image), svb::count_launch());
  else
    (image_sum_kernel<bf16><<<blocks, 256, 0, st>>>(static_cast<const bf16*>(t), out, n_images, per_image), svb::count_launch());
  SVB_LAUNCH_CHECK("image_sum");
  return 0;
}

extern "C" int svb_topk_columns(svb_handle* h, void* stream, const float* vals0, const int64_t* idx0,
                                const int64_t* files0, int32_t n0, const float* vals1, const int64_t* idx1,
                                const int64_t* files1, int32_t n1, int32_t F, int32_t k, int32_t largest,
                                float* out_vals, int64_t* out_idx, int64_t* out_files) {
  if (!h || !vals0 || !out_vals) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (n0 <= 0 || n1 < 0 || F <= 0 || k <= 0) return fail(SVB_ERR_BAD_ARG, "empty input");
  if (n1 > 0 && !vals1) return fail(SVB_ERR_BAD_ARG, "second source is null");
  const int n = n0 + n1;
  if (k > n) return fail(SVB_ERR_BAD_ARG, "k=%d exceeds the %d candidates per column", k, n);
  int P = 2;
  while (P < n) P <<= 1;
  if (P > 16384) return fail(SVB_ERR_UNSUPPORTED, "more than 16384 candidates per column");
  const size_t smem = static_cast<size_t>(P) * 8;
  if (smem > 48 * 1024) {
    static bool raised = false;
    if (!raised) {
      SVB_CUDA(cudaFuncSetAttribute(topk_columns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
      raised = true;
    }
  }
  (topk_columns_kernel<<<F, 256, smem, static_cast<cudaStream_t>(stream)>>>(
       vals0, idx0, files0, n0, vals1, idx1, files1, n1, F, k, largest, P, out_vals, out_idx, out_files),
   svb::count_launch());
  SVB_LAUNCH_CHECK("topk_columns");
  return 0;
}

extern "C" int svb_histogram_update(svb_handle* h, void* stream, const float* vals, int64_t rows, int32_t F,
                                    const int64_t* unit_idx, int32_t n_units, const float* mins, const float* maxs,
                                    int32_t bins, float* hist) {
  if (!h || !vals || !mins || !maxs || !hist) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (rows <= 0 || F <= 0 || n_units <= 0 || bins <= 0 || bins > 8192) return fail(SVB_ERR_BAD_ARG, "bad histogram shape");
  (histogram_columns_kernel<<<n_units, 256, static_cast<size_t>(bins) * 4, static_cast<cudaStream_t>(stream)>>>(
       vals, rows, F, unit_idx, mins, maxs, bins, hist, n_units),
   svb::count_launch());
  SVB_LAUNCH_CHECK("histogram_update");
  return 0;
}
