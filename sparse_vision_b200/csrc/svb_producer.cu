// The memory-bound layers of the frozen activation producer (the reference builds the
// producer at utils.py:277-281 and runs it around the hook at model_pipeline.py:445-475, 662-708).  torchvision's
// GoogLeNet in bf16 / channels_last spends two thirds of its forward in three kinds of eager kernels that only move
// bytes: ATen's NHWC max-pool (13 launches, ~0.7 TB/s), the broadcast bias add_ and relu_ behind every cuDNN
// convolution (2 x 57 read-modify-write passes) and the channel concatenation of every inception block.  Here:
//   * svb_maxpool_nhwc       -- 16-byte vectors over the channels, a strip of output columns per thread with the
//                               vertical maxima of every input column shared between the windows of the strip;
//   * svb_bias_relu_scatter  -- relu(conv + bias) in ONE pass, written straight into its channel range of the
//                               concatenated block output (or split over several destinations: the three 1x1
//                               convolutions on a block's input run as one convolution), so no cat pass exists.
// Both are exact: max is exact, and bf16(float(conv) + float(bias)) followed by max(., 0) is what add_ / relu_ compute.
#include "svb_common.cuh"
#include <cuda_bf16.h>

using namespace svb;

namespace {

__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
  // NaN-propagating like ATen's max_pool ((val > maxval) || isnan(val))
  __nv_bfloat162 r = __hmax2_nan(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ uint4 max8(uint4 a, uint4 b) {
  return make_uint4(max2(a.x, b.x), max2(a.y, b.y), max2(a.z, b.z), max2(a.w, b.w));
}

// in [N, H, W, C8] / out [N, OH, OW, C8] in 16-byte vectors of 8 bf16 channels.  One thread: one (image, output row,
// strip of SW output columns, channel vector); threads run over the channel vectors first (coalesced 16-byte accesses).
// The (SW-1)*S+K input columns of a strip are reduced vertically once and shared by the windows of the strip; all of
// a thread's loads are independent (18 / 27 of them in flight for 3x3 windows at stride 1 / 2).
template <int K, int S, int SW>
__global__ void __launch_bounds__(256)
maxpool_nhwc_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int H, int W, int C8, int pad, int OH,
                    int OW, int strips, long long total) {
  constexpr int NC = (SW - 1) * S + K;
  constexpr uint32_t kNegInf2 = 0xFF80FF80u;
  const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C8);
  long long r = i / C8;
  const int s = static_cast<int>(r % strips);
  r /= strips;
  const int oh = static_cast<int>(r % OH);
  const long long n = r / OH;
  const int ow0 = s * SW, ih0 = oh * S - pad, iw0 = ow0 * S - pad;
  const uint4* img = in + n * H * W * C8 + c;
  uint4 col[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    uint4 m = make_uint4(kNegInf2, kNegInf2, kNegInf2, kNegInf2);
    const int iw = iw0 + j;
    if (iw >= 0 && iw < W) {
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        const int ih = ih0 + kh;
        if (ih >= 0 && ih < H) m = max8(m, __ldg(img + (static_cast<long long>(ih) * W + iw) * C8));
      }
    }
    col[j] = m;
  }
  uint4* orow = out + ((n * OH + oh) * OW + ow0) * C8 + c;
#pragma unroll
  for (int o = 0; o < SW; ++o) {
    if (ow0 + o < OW) {
      uint4 m = col[o * S];
#pragma unroll
      for (int k = 1; k < K; ++k) m = max8(m, col[o * S + k]);
      orow[static_cast<long long>(o) * C8] = m;
    }
  }
}

struct SegTable {
  uint4* dst[SVB_MAX_CHAN_SEGMENTS];
  int begin8[SVB_MAX_CHAN_SEGMENTS], count8[SVB_MAX_CHAN_SEGMENTS], ld8[SVB_MAX_CHAN_SEGMENTS], off8[SVB_MAX_CHAN_SEGMENTS];
  int n;
};

__device__ __forceinline__ uint32_t bias_relu2(uint32_t v, uint32_t b, bool relu) {
  // add_ on bf16 tensors: both operands widened to fp32, one rounding; relu_ = clamp_min(0) on the rounded value
  float lo = __uint_as_float(v << 16) + __uint_as_float(b << 16);
  float hi = __uint_as_float(v & 0xFFFF0000u) + __uint_as_float(b & 0xFFFF0000u);
  __nv_bfloat162 r = __floats2bfloat162_rn(lo, hi);
  if (relu) r = __hmax2_nan(r, __floats2bfloat162_rn(0.f, 0.f));
  return *reinterpret_cast<uint32_t*>(&r);
}

// src [positions, C8] (dense convolution output), bias [C8]; every channel vector goes to the destination whose
// channel range holds it: dst[seg][(t * ld8 + off8 + (c - begin8))].  Two vectors per thread, half the tensor apart.
__device__ __forceinline__ void bias_relu_emit(long long i, uint4 v, const uint4* __restrict__ bias, const SegTable& tab,
                                               int C8, bool relu) {
  const int c = static_cast<int>(i % C8);
  const long long t = i / C8;
  const uint4 b = __ldg(bias + c);
  const uint4 o = make_uint4(bias_relu2(v.x, b.x, relu), bias_relu2(v.y, b.y, relu), bias_relu2(v.z, b.z, relu),
                             bias_relu2(v.w, b.w, relu));
#pragma unroll
  for (int sgi = 0; sgi < SVB_MAX_CHAN_SEGMENTS; ++sgi) {
    if (sgi < tab.n && c >= tab.begin8[sgi] && c < tab.begin8[sgi] + tab.count8[sgi])
      tab.dst[sgi][t * tab.ld8[sgi] + tab.off8[sgi] + (c - tab.begin8[sgi])] = o;
  }
}
__global__ void __launch_bounds__(256)
bias_relu_scatter_kernel(const uint4* __restrict__ src, const uint4* __restrict__ bias, const __grid_constant__ SegTable tab,
                         int C8, long long total, int relu) {
  const long long half = (total + 1) / 2;
  const long long i0 = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (i0 >= half) return;
  const long long i1 = i0 + half;
  const bool two = i1 < total;
  const uint4 v0 = __ldcs(src + i0);
  const uint4 v1 = two ? __ldcs(src + i1) : make_uint4(0, 0, 0, 0);
  bias_relu_emit(i0, v0, bias, tab, C8, relu != 0);
  if (two) bias_relu_emit(i1, v1, bias, tab, C8, relu != 0);
}

template <int K, int S>
void launch_pool(cudaStream_t st, const void* in, void* out, int64_t N, int H, int W, int C8, int pad, int OH, int OW) {
  constexpr int SW = 4;
  const int strips = (OW + SW - 1) / SW;
  const long long total = static_cast<long long>(N) * OH * strips * C8;
  (maxpool_nhwc_kernel<K, S, SW><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
       static_cast<const uint4*>(in), static_cast<uint4*>(out), H, W, C8, pad, OH, OW, strips, total),
   svb::count_launch());
}

// torch's pooling_output_shape (ATen/native/Pool.h) for dilation 1
int pool_out_size(int in, int k, int s, int pad, int ceil_mode) {
  int o = (in + 2 * pad - (k - 1) - 1 + (ceil_mode ? s - 1 : 0)) / s + 1;
  if (ceil_mode && (o - 1) * s >= in + pad) --o;
  return o;
}

}  // namespace

extern "C" int svb_maxpool_nhwc(svb_handle* h, void* stream, const void* in, int64_t n_images, int32_t H, int32_t W,
                                int32_t C, int32_t kernel, int32_t stride, int32_t pad, int32_t ceil_mode, void* out,
                                int32_t OH, int32_t OW) {
  if (!h || !in || !out) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (n_images <= 0 || H <= 0 || W <= 0 || C <= 0) return fail(SVB_ERR_BAD_ARG, "empty tensor");
  if (C % 8 || (reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return fail(SVB_ERR_UNSUPPORTED, "svb_maxpool_nhwc needs C %% 8 == 0 (C = %d) and 16-byte aligned pointers", C);
  if (pad < 0 || 2 * pad > kernel) return fail(SVB_ERR_BAD_ARG, "pad must be at most half the kernel size");
  if (OH != pool_out_size(H, kernel, stride, pad, ceil_mode) || OW != pool_out_size(W, kernel, stride, pad, ceil_mode))
    return fail(SVB_ERR_BAD_ARG, "output is %d x %d, expected %d x %d", OH, OW,
                pool_out_size(H, kernel, stride, pad, ceil_mode), pool_out_size(W, kernel, stride, pad, ceil_mode));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C8 = C / 8;
  if (kernel == 3 && stride == 1) launch_pool<3, 1>(st, in, out, n_images, H, W, C8, pad, OH, OW);
  else if (kernel == 3 && stride == 2) launch_pool<3, 2>(st, in, out, n_images, H, W, C8, pad, OH, OW);
  else if (kernel == 2 && stride == 2) launch_pool<2, 2>(st, in, out, n_images, H, W, C8, pad, OH, OW);
  else return fail(SVB_ERR_UNSUPPORTED, "max-pool %dx%d stride %d is not one of GoogLeNet's (3/1, 3/2, 2/2)", kernel,
                   kernel, stride);
  SVB_LAUNCH_CHECK("maxpool_nhwc");
  return 0;
}

extern "C" int svb_bias_relu_scatter(svb_handle* h, void* stream, const void* src, const void* bias, int64_t positions,
                                     int32_t C, const svb_chan_segment* seg, int32_t n_seg, int32_t relu) {
  if (!h || !src || !bias || !seg) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (positions <= 0 || C <= 0) return fail(SVB_ERR_BAD_ARG, "empty tensor");
  if (n_seg < 1 || n_seg > SVB_MAX_CHAN_SEGMENTS) return fail(SVB_ERR_BAD_ARG, "1..%d destinations", SVB_MAX_CHAN_SEGMENTS);
  if (C % 8 || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(bias) & 15))
    return fail(SVB_ERR_UNSUPPORTED, "svb_bias_relu_scatter needs C %% 8 == 0 (C = %d) and 16-byte aligned pointers", C);
  SegTable tab;
  tab.n = n_seg;
  int covered = 0;
  for (int i = 0; i < n_seg; ++i) {
    const svb_chan_segment& s = seg[i];
    if (!s.dst || (reinterpret_cast<uintptr_t>(s.dst) & 15) || s.c_begin % 8 || s.c_count % 8 || s.dst_channels % 8 ||
        s.dst_offset % 8 || s.c_count <= 0 || s.c_begin != covered || s.dst_offset < 0 ||
        s.dst_offset + s.c_count > s.dst_channels)
      return fail(SVB_ERR_BAD_ARG, "destination %d: channel ranges must be multiples of 8, consecutive from 0 and inside "
                  "the destination's row", i);
    covered += s.c_count;
    tab.dst[i] = static_cast<uint4*>(s.dst);
    tab.begin8[i] = s.c_begin / 8; tab.count8[i] = s.c_count / 8; tab.ld8[i] = s.dst_channels / 8; tab.off8[i] = s.dst_offset / 8;
  }
  if (covered != C) return fail(SVB_ERR_BAD_ARG, "the destinations cover %d of %d channels", covered, C);
  const long long total = positions * (C / 8);
  const long long half = (total + 1) / 2;
  (bias_relu_scatter_kernel<<<static_cast<unsigned>((half + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
       static_cast<const uint4*>(src), static_cast<const uint4*>(bias), tab, C / 8, total, relu),
   svb::count_launch());
  SVB_LAUNCH_CHECK("bias_relu_scatter");
  return 0;
}
