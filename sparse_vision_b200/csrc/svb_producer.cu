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
// a thread's loads are independent and unpredicated (18 / 27 of them in flight for 3x3 windows at stride 1 / 2).
// Idx: unsigned (32-bit divisions; every GoogLeNet shape) or long long (more than 2^31 work items).
template <int K, int S, int SW, typename Idx>
__global__ void __launch_bounds__(256)
maxpool_nhwc_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int H, int W, int C8, int pad, int OH,
                    int OW, int strips, long long total) {
  constexpr int NC = (SW - 1) * S + K;
  const Idx i = static_cast<Idx>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= static_cast<Idx>(total)) return;
  const int c = static_cast<int>(i % static_cast<Idx>(C8));
  Idx r = i / static_cast<Idx>(C8);
  const int s = static_cast<int>(r % static_cast<Idx>(strips));
  r /= static_cast<Idx>(strips);
  const int oh = static_cast<int>(r % static_cast<Idx>(OH));
  const long long n = static_cast<long long>(r / static_cast<Idx>(OH));
  const int ow0 = s * SW, ih0 = oh * S - pad, iw0 = ow0 * S - pad;
  const uint4* img = in + n * H * W * C8 + c;
  // Positions outside the image count as -inf.  Every window that is stored starts inside the image (or its left / top
  // padding), so clamping a coordinate into the image only repeats a value the same window already holds: the loads
  // need no predicates and all of them are issued before the first max.
  int roff[K];
#pragma unroll
  for (int kh = 0; kh < K; ++kh) roff[kh] = min(max(ih0 + kh, 0), H - 1) * W;
  uint4 col[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int iw = min(max(iw0 + j, 0), W - 1);
    uint4 v[K];
#pragma unroll
    for (int kh = 0; kh < K; ++kh) v[kh] = __ldg(img + static_cast<long long>(roff[kh] + iw) * C8);
    uint4 m = v[0];
#pragma unroll
    for (int kh = 1; kh < K; ++kh) m = max8(m, v[kh]);
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
template <typename Idx>
__device__ __forceinline__ void bias_relu_emit(Idx i, uint4 v, const uint4* __restrict__ bias, const SegTable& tab,
                                               int C8, bool relu) {
  const int c = static_cast<int>(i % static_cast<Idx>(C8));
  const long long t = static_cast<long long>(i / static_cast<Idx>(C8));
  const uint4 b = __ldg(bias + c);
  const uint4 o = make_uint4(bias_relu2(v.x, b.x, relu), bias_relu2(v.y, b.y, relu), bias_relu2(v.z, b.z, relu),
                             bias_relu2(v.w, b.w, relu));
#pragma unroll
  for (int sgi = 0; sgi < SVB_MAX_CHAN_SEGMENTS; ++sgi) {
    if (sgi < tab.n && c >= tab.begin8[sgi] && c < tab.begin8[sgi] + tab.count8[sgi])
      tab.dst[sgi][t * tab.ld8[sgi] + tab.off8[sgi] + (c - tab.begin8[sgi])] = o;
  }
}
template <typename Idx>
__global__ void __launch_bounds__(256)
bias_relu_scatter_kernel(const uint4* __restrict__ src, const uint4* __restrict__ bias, const __grid_constant__ SegTable tab,
                         int C8, long long total_, int relu) {
  const Idx total = static_cast<Idx>(total_);
  const Idx half = (total + 1) / 2;
  const Idx i0 = static_cast<Idx>(blockIdx.x) * 256 + threadIdx.x;
  if (i0 >= half) return;
  const Idx i1 = i0 + half;
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
  if (total < (1LL << 31))
    (maxpool_nhwc_kernel<K, S, SW, unsigned><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
         static_cast<const uint4*>(in), static_cast<uint4*>(out), H, W, C8, pad, OH, OW, strips, total),
     svb::count_launch());
  else
    (maxpool_nhwc_kernel<K, S, SW, long long><<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
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
  if (total < (1LL << 31))
    (bias_relu_scatter_kernel<unsigned><<<static_cast<unsigned>((half + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
         static_cast<const uint4*>(src), static_cast<const uint4*>(bias), tab, C / 8, total, relu),
     svb::count_launch());
  else
    (bias_relu_scatter_kernel<long long><<<static_cast<unsigned>((half + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
         static_cast<const uint4*>(src), static_cast<const uint4*>(bias), tab, C / 8, total, relu),
     svb::count_launch());
  SVB_LAUNCH_CHECK("bias_relu_scatter");
  return 0;
}

namespace {
// ------------------------------------------------------------------------------------------------ differentiable max-pool
// The IE passes (compute_ie.py:270-311, get_grad_original) back-propagate the loss through the layers behind the first
// hooked one; ATen's max-pool forward + backward are 2.5 of the 5.5 ms of that pass (64 images).  Forward with indices:
// one thread per (image, output position, 8-channel vector) scans the window in ATen's order (kh outer, kw inner) with
// ATen's rule `(val > max) || isnan(val)` on packed bf16 pairs, so ties go to the FIRST maximum exactly as in
// max_pool2d_with_indices; the index is the window offset kh*K + kw (one byte per element).  All loads are issued
// unpredicated with clamped coordinates; positions outside the image are skipped in the comparison.
template <int K, int S>
__global__ void __launch_bounds__(256)
maxpool_nhwc_argmax_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, uint2* __restrict__ arg, int H, int W,
                           int C8, int pad, int OH, int OW, unsigned total) {
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % static_cast<unsigned>(C8));
  unsigned r = i / static_cast<unsigned>(C8);
  const int ow = static_cast<int>(r % static_cast<unsigned>(OW));
  r /= static_cast<unsigned>(OW);
  const int oh = static_cast<int>(r % static_cast<unsigned>(OH));
  const long long n = r / static_cast<unsigned>(OH);
  const int ih0 = oh * S - pad, iw0 = ow * S - pad;
  const uint4* img = in + n * H * W * C8 + c;
  uint4 v[K * K];
#pragma unroll
  for (int kh = 0; kh < K; ++kh)
#pragma unroll
    for (int kw = 0; kw < K; ++kw)
      v[kh * K + kw] = __ldg(img + (static_cast<long long>(min(max(ih0 + kh, 0), H - 1)) * W + min(max(iw0 + kw, 0), W - 1)) * C8);
  uint32_t m[4] = {0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u};   // -inf pairs
  uint32_t idx[4];
  bool first = true;
#pragma unroll
  for (int kh = 0; kh < K; ++kh) {
#pragma unroll
    for (int kw = 0; kw < K; ++kw) {
      const int ih = ih0 + kh, iw = iw0 + kw;
      if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
      const uint32_t code = static_cast<uint32_t>(kh * K + kw) * 0x00010001u;
      if (first) {   // ATen starts maxidx at the first position inside the image
        first = false;
#pragma unroll
        for (int q = 0; q < 4; ++q) idx[q] = code;
      }
      const uint32_t w[4] = {v[kh * K + kw].x, v[kh * K + kw].y, v[kh * K + kw].z, v[kh * K + kw].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&m[q]);
        // (val > max) || isnan(val): ordered greater-than, or val unordered with itself
        const uint32_t take = __hgt2_mask(a, b) | __hneu2_mask(a, a);
        m[q] = (w[q] & take) | (m[q] & ~take);
        idx[q] = (code & take) | (idx[q] & ~take);
      }
    }
  }
  const long long o = ((n * OH + oh) * OW + ow) * C8 + c;
  out[o] = make_uint4(m[0], m[1], m[2], m[3]);
  // pack the eight 16-bit indices into eight bytes (element order preserved)
  uint2 a8;
  a8.x = (idx[0] & 0xFFu) | ((idx[0] >> 8) & 0xFF00u) | ((idx[1] & 0xFFu) << 16) | ((idx[1] & 0xFF0000u) << 8);
  a8.y = (idx[2] & 0xFFu) | ((idx[2] >> 8) & 0xFF00u) | ((idx[3] & 0xFFu) << 16) | ((idx[3] & 0xFF0000u) << 8);
  arg[o] = a8;
}

// grad_in[n, ih, iw, c] = sum of grad_out over the windows whose maximum sits at (ih, iw): a gather over the at most
// ceil(K/S)^2 windows that cover the position (deterministic, no atomics), fp32 accumulation, one rounding.
template <int K, int S>
__global__ void __launch_bounds__(256)
maxpool_nhwc_backward_kernel(const uint4* __restrict__ gout, const uint2* __restrict__ arg, uint4* __restrict__ gin,
                             int H, int W, int C8, int pad, int OH, int OW, unsigned total) {
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % static_cast<unsigned>(C8));
  unsigned r = i / static_cast<unsigned>(C8);
  const int iw = static_cast<int>(r % static_cast<unsigned>(W));
  r /= static_cast<unsigned>(W);
  const int ih = static_cast<int>(r % static_cast<unsigned>(H));
  const long long n = r / static_cast<unsigned>(H);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
  for (int kh = 0; kh < K; ++kh) {
    const int th = ih + pad - kh;
    if (th < 0 || th % S != 0 || th / S >= OH) continue;
#pragma unroll
    for (int kw = 0; kw < K; ++kw) {
      const int tw = iw + pad - kw;
      if (tw < 0 || tw % S != 0 || tw / S >= OW) continue;
      const long long o = ((n * OH + th / S) * OW + tw / S) * C8 + c;
      const uint2 a8 = __ldg(arg + o);
      const uint4 g = __ldg(gout + o);
      const uint32_t code = static_cast<uint32_t>(kh * K + kw);
      const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t byte = ((e < 4 ? a8.x : a8.y) >> (8 * (e & 3))) & 0xFFu;
        const float ge = __uint_as_float((e & 1) ? (gw[e >> 1] & 0xFFFF0000u) : (gw[e >> 1] << 16));
        if (byte == code) acc[e] += ge;
      }
    }
  }
  uint32_t o4[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    __nv_bfloat162 p2 = __floats2bfloat162_rn(acc[2 * q], acc[2 * q + 1]);
    o4[q] = *reinterpret_cast<uint32_t*>(&p2);
  }
  gin[i] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
}
}  // namespace

extern "C" int svb_maxpool_nhwc_argmax(svb_handle* h, void* stream, const void* in, int64_t n_images, int32_t H, int32_t W,
                                       int32_t C, int32_t kernel, int32_t stride, int32_t pad, int32_t ceil_mode,
                                       void* out, uint8_t* argmax, int32_t OH, int32_t OW) {
  if (!h || !in || !out || !argmax) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (n_images <= 0 || H <= 0 || W <= 0 || C <= 0) return fail(SVB_ERR_BAD_ARG, "empty tensor");
  if (C % 8 || (reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(argmax) & 7))
    return fail(SVB_ERR_UNSUPPORTED, "svb_maxpool_nhwc_argmax needs C %% 8 == 0 (C = %d) and aligned pointers", C);
  if (pad < 0 || 2 * pad > kernel) return fail(SVB_ERR_BAD_ARG, "pad must be at most half the kernel size");
  if (OH != pool_out_size(H, kernel, stride, pad, ceil_mode) || OW != pool_out_size(W, kernel, stride, pad, ceil_mode))
    return fail(SVB_ERR_BAD_ARG, "output is %d x %d, expected %d x %d", OH, OW,
                pool_out_size(H, kernel, stride, pad, ceil_mode), pool_out_size(W, kernel, stride, pad, ceil_mode));
  const long long total = n_images * OH * OW * (C / 8);
  if (total >= (1LL << 32) || n_images * H * W * (C / 8) >= (1LL << 32)) return fail(SVB_ERR_UNSUPPORTED, "tensor too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256), tot = static_cast<unsigned>(total);
#define SVB_POOL_ARG(K_, S_)                                                                                           \
  (maxpool_nhwc_argmax_kernel<K_, S_><<<blocks, 256, 0, st>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), \
                                                               reinterpret_cast<uint2*>(argmax), H, W, C / 8, pad, OH, \
                                                               OW, tot),                                               \
   svb::count_launch())
  if (kernel == 3 && stride == 1) SVB_POOL_ARG(3, 1);
  else if (kernel == 3 && stride == 2) SVB_POOL_ARG(3, 2);
  else if (kernel == 2 && stride == 2) SVB_POOL_ARG(2, 2);
  else return fail(SVB_ERR_UNSUPPORTED, "max-pool %dx%d stride %d is not one of GoogLeNet's (3/1, 3/2, 2/2)", kernel,
                   kernel, stride);
#undef SVB_POOL_ARG
  SVB_LAUNCH_CHECK("maxpool_nhwc_argmax");
  return 0;
}

extern "C" int svb_maxpool_nhwc_backward(svb_handle* h, void* stream, const void* grad_out, const uint8_t* argmax,
                                         int64_t n_images, int32_t H, int32_t W, int32_t C, int32_t kernel,
                                         int32_t stride, int32_t pad, int32_t OH, int32_t OW, void* grad_in) {
  if (!h || !grad_out || !argmax || !grad_in) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (n_images <= 0 || H <= 0 || W <= 0 || C <= 0 || OH <= 0 || OW <= 0) return fail(SVB_ERR_BAD_ARG, "empty tensor");
  if (C % 8 || (reinterpret_cast<uintptr_t>(grad_out) & 15) || (reinterpret_cast<uintptr_t>(grad_in) & 15) ||
      (reinterpret_cast<uintptr_t>(argmax) & 7))
    return fail(SVB_ERR_UNSUPPORTED, "svb_maxpool_nhwc_backward needs C %% 8 == 0 (C = %d) and aligned pointers", C);
  const long long total = n_images * H * W * (C / 8);
  if (total >= (1LL << 32) || n_images * OH * OW * (C / 8) >= (1LL << 32)) return fail(SVB_ERR_UNSUPPORTED, "tensor too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256), tot = static_cast<unsigned>(total);
#define SVB_POOL_BWD(K_, S_)                                                                                       \
  (maxpool_nhwc_backward_kernel<K_, S_><<<blocks, 256, 0, st>>>(static_cast<const uint4*>(grad_out),               \
                                                                 reinterpret_cast<const uint2*>(argmax),            \
                                                                 static_cast<uint4*>(grad_in), H, W, C / 8, pad, OH, \
                                                                 OW, tot),                                          \
   svb::count_launch())
  if (kernel == 3 && stride == 1) SVB_POOL_BWD(3, 1);
  else if (kernel == 3 && stride == 2) SVB_POOL_BWD(3, 2);
  else if (kernel == 2 && stride == 2) SVB_POOL_BWD(2, 2);
  else return fail(SVB_ERR_UNSUPPORTED, "max-pool %dx%d stride %d is not one of GoogLeNet's (3/1, 3/2, 2/2)", kernel,
                   kernel, stride);
#undef SVB_POOL_BWD
  SVB_LAUNCH_CHECK("maxpool_nhwc_backward");
  return 0;
}

namespace {
// ------------------------------------------------------------------------------------------------ backward of bias + relu + concat
// The inverse of bias_relu_scatter for the IE passes: dst[t, begin + c] = y[t, y_off + c] > 0 ? g[t, g_off + c] : 0, i.e.
// relu's backward (threshold_backward: grad * (result > 0)) reading the gradient of a block output -- or of a 3x3-reduce
// tensor -- at its channel range and writing the DENSE gradient of the convolution output that produced the range.
struct GatherTable {
  const uint4* g[SVB_MAX_CHAN_SEGMENTS];
  const uint4* y[SVB_MAX_CHAN_SEGMENTS];
  int begin8[SVB_MAX_CHAN_SEGMENTS], count8[SVB_MAX_CHAN_SEGMENTS];
  int g_ld8[SVB_MAX_CHAN_SEGMENTS], g_off8[SVB_MAX_CHAN_SEGMENTS], y_ld8[SVB_MAX_CHAN_SEGMENTS], y_off8[SVB_MAX_CHAN_SEGMENTS];
  int n;
};
__device__ __forceinline__ uint32_t relu_grad2(uint32_t g, uint32_t y) {
  const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
  return g & __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&y), zero);
}
__global__ void __launch_bounds__(256)
relu_grad_gather_kernel(uint4* __restrict__ dst, const __grid_constant__ GatherTable tab, int C8, unsigned total) {
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % static_cast<unsigned>(C8));
  const long long t = i / static_cast<unsigned>(C8);
  uint4 o = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int s = 0; s < SVB_MAX_CHAN_SEGMENTS; ++s) {
    if (s < tab.n && c >= tab.begin8[s] && c < tab.begin8[s] + tab.count8[s]) {
      const int cc = c - tab.begin8[s];
      const uint4 g = __ldg(tab.g[s] + t * tab.g_ld8[s] + tab.g_off8[s] + cc);
      const uint4 y = __ldg(tab.y[s] + t * tab.y_ld8[s] + tab.y_off8[s] + cc);
      o = make_uint4(relu_grad2(g.x, y.x), relu_grad2(g.y, y.y), relu_grad2(g.z, y.z), relu_grad2(g.w, y.w));
    }
  }
  dst[i] = o;
}
}  // namespace

extern "C" int svb_relu_grad_gather(svb_handle* h, void* stream, int64_t positions, int32_t C,
                                    const svb_grad_segment* seg, int32_t n_seg, void* dst) {
  if (!h || !seg || !dst) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (positions <= 0 || C <= 0) return fail(SVB_ERR_BAD_ARG, "empty tensor");
  if (n_seg < 1 || n_seg > SVB_MAX_CHAN_SEGMENTS) return fail(SVB_ERR_BAD_ARG, "1..%d sources", SVB_MAX_CHAN_SEGMENTS);
  if (C % 8 || (reinterpret_cast<uintptr_t>(dst) & 15))
    return fail(SVB_ERR_UNSUPPORTED, "svb_relu_grad_gather needs C %% 8 == 0 (C = %d) and a 16-byte aligned destination", C);
  GatherTable tab;
  tab.n = n_seg;
  int covered = 0;
  for (int i = 0; i < n_seg; ++i) {
    const svb_grad_segment& s = seg[i];
    if (!s.grad || !s.y || (reinterpret_cast<uintptr_t>(s.grad) & 15) || (reinterpret_cast<uintptr_t>(s.y) & 15) ||
        s.c_begin != covered || s.c_count <= 0 || s.c_count % 8 || s.grad_channels % 8 || s.grad_offset % 8 ||
        s.y_channels % 8 || s.y_offset % 8 || s.grad_offset < 0 || s.y_offset < 0 ||
        s.grad_offset + s.c_count > s.grad_channels || s.y_offset + s.c_count > s.y_channels)
      return fail(SVB_ERR_BAD_ARG, "source %d: channel ranges must be multiples of 8, consecutive from 0 and inside the "
                  "source rows", i);
    covered += s.c_count;
    tab.g[i] = static_cast<const uint4*>(s.grad); tab.y[i] = static_cast<const uint4*>(s.y);
    tab.begin8[i] = s.c_begin / 8; tab.count8[i] = s.c_count / 8;
    tab.g_ld8[i] = s.grad_channels / 8; tab.g_off8[i] = s.grad_offset / 8;
    tab.y_ld8[i] = s.y_channels / 8; tab.y_off8[i] = s.y_offset / 8;
  }
  if (covered != C) return fail(SVB_ERR_BAD_ARG, "the sources cover %d of %d channels", covered, C);
  const long long total = positions * (C / 8);
  if (total >= (1LL << 32)) return fail(SVB_ERR_UNSUPPORTED, "tensor too large");
  (relu_grad_gather_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
       static_cast<uint4*>(dst), tab, C / 8, static_cast<unsigned>(total)),
   svb::count_launch());
  SVB_LAUNCH_CHECK("relu_grad_gather");
  return 0;
}

namespace {
// ------------------------------------------------------------------------------------------------ stem convolution
// GoogLeNet's conv1 (7x7, stride 2, pad 3, 3 -> 64 channels, 224x224 -> 112x112; googlenet.py of torchvision, built by
// the reference at utils.py:277-281) + folded BatchNorm bias + ReLU.  cuDNN pads the 3 input channels to 8 and runs a
// 256x64 implicit-GEMM kernel: 1.6 ms for 256 images, 40 % of the whole fused forward.  With 3 channels the 7 kernel
// columns of one kernel row are 21 CONTIGUOUS bf16 values of the NHWC input row, so the im2col matrix never has to
// exist: for kernel row kh,  A[ow][kk] = in_row[2*oh + kh - 3][6*ow + kk - 10]  for kk = 1 + 3*kw + c  (kk = 0 meets a
// zero weight), i.e. the A fragments of mma.sync.m16n8k16 are plain 32-bit shared-memory loads from the staged input
// rows at a 12-byte row pitch.  The one-element shift (kk = 1 + ...) makes every fragment pair 4-byte aligned while the
// global rows are staged with aligned 16-byte copies.  K is FLATTENED over the kernel rows: kf = 22*kh + kk, 154 values
// padded to 160 = ten k16 steps (padding each kernel row to 32 took fourteen); a fragment pair never straddles a kernel
// row (22 is even), a k-step may, which costs a two-way bank conflict on those loads.
// CTA = 4 output rows of one image (448 pixels) x 64 channels, 7 warps, each warp two passes of 32 pixels x 64
// channels (64 fp32 accumulators).  Shared memory: 13 input rows (zero padded) + the weights as [64][168] bf16 (pitch
// 336 B: the eight rows of an ldmatrix fall on distinct banks), 39 808 B static.
constexpr int kC1RowsPerCta = 4, kC1InRows = 2 * kC1RowsPerCta + 5, kC1RowElems = 704, kC1WPitch = 168;
constexpr int kC1KSteps = 10, kC1KRow = 22, kC1K = 7 * kC1KRow;
constexpr int kC1Threads = 224;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_row));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// 16-byte global -> shared copy that bypasses the registers; src_bytes = 0 writes zeros (the zero padding of the rows)
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kC1InBytes = kC1InRows * kC1RowElems * 2;                   // one staged tile of input rows
constexpr int kC1OutPitch = 144, kC1OutBytes = 16 * kC1OutPitch;          // per-warp output staging: 16 pixels x (128 + 16) B
constexpr int kC1SmemBytes = 64 * kC1WPitch * 2 + 2 * kC1InBytes + (kC1Threads / 32) * kC1OutBytes;   // 74 240 B

// Persistent: the weights are staged once per CTA, the CTA walks tiles (image, block of 4 output rows) with a stride of
// the grid, and the input rows of the NEXT tile arrive by cp.async while the current one is multiplied.
__global__ void __launch_bounds__(kC1Threads, 2)
conv1_7x7s2_kernel(const uint4* __restrict__ x, const uint4* __restrict__ wt, const __nv_bfloat16* __restrict__ bias,
                   __nv_bfloat16* __restrict__ out, int relu, int n_tiles) {
  extern __shared__ __align__(16) uint8_t c1_smem[];
  uint16_t* s_w = reinterpret_cast<uint16_t*>(c1_smem);
  uint8_t* s_in_base = c1_smem + 64 * kC1WPitch * 2;
  const int tid = threadIdx.x;
  constexpr int kVecPerRow = kC1RowElems / 8;          // 88: 2 zero vectors, 84 of data (224 px x 3 ch), 2 zero vectors
  auto stage = [&](int tile, int buf) {
    const int n = tile / 28, oh0 = (tile % 28) * kC1RowsPerCta;
    uint4* dst = reinterpret_cast<uint4*>(s_in_base + buf * kC1InBytes);
    for (int i = tid; i < kC1InRows * kVecPerRow; i += kC1Threads) {
      const int r = i / kVecPerRow, v = i % kVecPerRow, ih = 2 * oh0 - 3 + r;
      const bool valid = v >= 2 && v < 86 && ih >= 0 && ih < 224;
      const uint4* src = valid ? x + (static_cast<long long>(n) * 224 + ih) * 84 + (v - 2) : x;
      cp_async16_zfill(dst + i, src, valid ? 16 : 0);
    }
  };
  int tile = blockIdx.x, buf = 0;
  if (tile < n_tiles) stage(tile, 0);
  cp_async_commit();
  for (int i = tid; i < 64 * kC1WPitch / 8; i += kC1Threads) reinterpret_cast<uint4*>(s_w)[i] = __ldg(wt + i);
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  // ldmatrix.x4: lane l addresses row (l & 7) of matrix (l >> 3); matrices = (n-tile, k 0-7), (n-tile, k 8-15),
  // (n-tile + 1, k 0-7), (n-tile + 1, k 8-15)
  const uint16_t* b_lane = s_w + ((lane & 7) + ((lane >> 4) << 3)) * kC1WPitch + (((lane >> 3) & 1) << 3);
  uint8_t* s_out = s_in_base + 2 * kC1InBytes + warp * kC1OutBytes;
  for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
    const int next = tile + gridDim.x;
    if (next < n_tiles) stage(next, buf ^ 1);       // buf ^ 1 was released by the barrier at the end of the last round
    cp_async_commit();
    cp_async_wait<1>();                             // this tile's rows have landed (the next tile's may be in flight)
    __syncthreads();
    const uint32_t* s_in32 = reinterpret_cast<const uint32_t*>(s_in_base + buf * kC1InBytes);
    const int n = tile / 28, oh0 = (tile % 28) * kC1RowsPerCta;
  #pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int mt0 = 4 * warp + 2 * pass;
      int a_base[2], rl[2], owb[2];
  #pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        rl[mi] = (mt0 + mi) / 7;
        owb[mi] = ((mt0 + mi) % 7) * 16;
        a_base[mi] = (2 * rl[mi]) * (kC1RowElems / 2) + 3 * (owb[mi] + g) + 3;   // word index of (kernel row 0, kk = 0)
      }
      float acc[2][8][4];
  #pragma unroll
      for (int mi = 0; mi < 2; ++mi)
  #pragma unroll
        for (int nt = 0; nt < 8; ++nt)
  #pragma unroll
          for (int q = 0; q < 4; ++q) acc[mi][nt][q] = 0.f;
  #pragma unroll
      for (int j = 0; j < kC1KSteps; ++j) {
        uint32_t b[4][4];
  #pragma unroll
        for (int np = 0; np < 4; ++np) ldmatrix_x4(b[np], b_lane + np * 16 * kC1WPitch + j * 16);
        // word offsets of this thread's two fragment pairs (kf = 16 j + 2 t and + 8): kernel row kh = kf / 22 (the six
        // padding values behind kf = 153 stay in kernel row 6 and meet zero weights), kk = kf - 22 kh
        int off[2];
  #pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int kf = 16 * j + 8 * hh + 2 * t;
          const int kh = min((kf * 373) >> 13, 6);
          off[hh] = kh * (kC1RowElems / 2) + ((kf - kC1KRow * kh) >> 1);
        }
  #pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          uint32_t a[4];
          a[0] = s_in32[a_base[mi] + off[0]];
          a[1] = s_in32[a_base[mi] + off[0] + 24];
          a[2] = s_in32[a_base[mi] + off[1]];
          a[3] = s_in32[a_base[mi] + off[1] + 24];
  #pragma unroll
          for (int np = 0; np < 4; ++np) {
            mma_bf16_16816(acc[mi][2 * np], a, b[np][0], b[np][1]);
            mma_bf16_16816(acc[mi][2 * np + 1], a, b[np][2], b[np][3]);
          }
        }
      }
      // Epilogue: bias + relu, then the warp's 16 pixels x 64 channels go through a padded shared-memory tile (stmatrix:
      // the inverse of ldmatrix; 144-byte row pitch keeps the eight 16-byte rows of a matrix on distinct banks) so that
      // every global store is 16 bytes per lane and 128 contiguous bytes per pixel.  (Storing the accumulator fragments
      // directly -- 4 bytes per lane, eight 16-byte pieces per instruction -- cost as many L1 wavefronts as all the
      // fragment loads of the kernel.)
  #pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        __syncwarp();
  #pragma unroll
        for (int r = 0; r < 2; ++r) {
  #pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t pk[4];
  #pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const int nt = 4 * q + i4;
              const __nv_bfloat162 bb = *reinterpret_cast<const __nv_bfloat162*>(bias + nt * 8 + 2 * t);
              float v0 = acc[mi][nt][2 * r] + __low2float(bb), v1 = acc[mi][nt][2 * r + 1] + __high2float(bb);
              if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
              const __nv_bfloat162 o2 = __floats2bfloat162_rn(v0, v1);
              pk[i4] = *reinterpret_cast<const uint32_t*>(&o2);
            }
            // lane l addresses row (l & 7) of matrix (l >> 3) = n-tile 4 q + (l >> 3)
            const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(
                s_out + (8 * r + (lane & 7)) * kC1OutPitch + (4 * q + (lane >> 3)) * 16));
            asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk[0]),
                         "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
          }
        }
        __syncwarp();
        uint4* opix = reinterpret_cast<uint4*>(out + ((static_cast<long long>(n) * 112 + oh0 + rl[mi]) * 112 + owb[mi]) * 64);
  #pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int c = it * 32 + lane, row = c >> 3, col = c & 7;
          opix[row * 8 + col] = *reinterpret_cast<const uint4*>(s_out + row * kC1OutPitch + col * 16);
        }
      }
    }
    __syncthreads();                                // every warp is done with this buffer before it is staged again
  }
}

// w [64, 3, 7, 7] (any strides, bf16) -> wt [64][168]: wt[n][22 * kh + 1 + 3 * kw + c] = w[n][c][kh][kw], zeros elsewhere
__global__ void conv1_pack_weights_kernel(const __nv_bfloat16* __restrict__ w, long long sn, long long sc, long long sh,
                                          long long sw, uint16_t* __restrict__ wt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * kC1WPitch) return;
  const int n = i / kC1WPitch, k = i % kC1WPitch, kh = k / kC1KRow, kk = k % kC1KRow - 1;
  uint16_t v = 0;
  if (k < kC1K && kk >= 0)
    v = *reinterpret_cast<const uint16_t*>(w + n * sn + (kk % 3) * sc + kh * sh + (kk / 3) * sw);
  wt[i] = v;
}
}  // namespace

extern "C" int svb_conv1_pack_weights(svb_handle* h, void* stream, const void* w, int64_t stride_o, int64_t stride_i,
                                      int64_t stride_h, int64_t stride_w, void* packed) {
  if (!h || !w || !packed) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (reinterpret_cast<uintptr_t>(packed) & 15) return fail(SVB_ERR_BAD_ARG, "packed weights must be 16-byte aligned");
  (conv1_pack_weights_kernel<<<cdiv(64 * kC1WPitch, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
       static_cast<const __nv_bfloat16*>(w), stride_o, stride_i, stride_h, stride_w, static_cast<uint16_t*>(packed)),
   svb::count_launch());
  SVB_LAUNCH_CHECK("conv1_pack_weights");
  return 0;
}

extern "C" int svb_conv1_7x7s2_nhwc(svb_handle* h, void* stream, const void* x, int64_t n_images, const void* packed_w,
                                    const void* bias, int32_t relu, void* out) {
  if (!h || !x || !packed_w || !bias || !out) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (n_images <= 0 || n_images * 28 > 2147483647LL) return fail(SVB_ERR_BAD_ARG, "bad image count");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(packed_w) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(bias) & 3))
    return fail(SVB_ERR_UNSUPPORTED, "svb_conv1_7x7s2_nhwc needs 16-byte aligned tensors");
  static bool configured[64] = {};
  if (h->device < 64 && !configured[h->device]) {
    SVB_CUDA(cudaFuncSetAttribute(conv1_7x7s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kC1SmemBytes));
    configured[h->device] = true;
  }
  const int n_tiles = static_cast<int>(n_images * 28);
  const int grid = n_tiles < 2 * h->sms ? n_tiles : 2 * h->sms;       // two resident CTAs per SM walk the tiles
  (conv1_7x7s2_kernel<<<grid, kC1Threads, kC1SmemBytes, static_cast<cudaStream_t>(stream)>>>(
       static_cast<const uint4*>(x), static_cast<const uint4*>(packed_w), static_cast<const __nv_bfloat16*>(bias),
       static_cast<__nv_bfloat16*>(out), relu, n_tiles),
   svb::count_launch());
  SVB_LAUNCH_CHECK("conv1_7x7s2");
  return 0;
}
