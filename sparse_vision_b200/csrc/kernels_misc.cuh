// Memory-bound helper kernels around the GEMMs: layout packing, weight preparation, deterministic reductions,
// per-channel statistics, the fused (Constrained)Adam update, activity finalisation and the step scalars.
// All reductions use a fixed summation order (no floating-point atomics) so a step is reproducible.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "ptx.cuh"

namespace svb {

typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float block_sum(float v, float* smem /* >= 32 floats */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = lane < nw ? smem[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;  // valid in warp 0
}

// ------------------------------------------------------------------------------------------------ layout
// tokens [B*HW, C] -> NCHW [B,C,HW]; utils.py:2478 '(b h w) c -> b c h w'.
template <typename TIn, typename TOut>
static __global__ void unpack_tokens_to_nchw_kernel(const TIn* __restrict__ tok, TOut* __restrict__ out, int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const TIn* tb = tok + static_cast<size_t>(b) * HW * C;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < HW && c < C) ? to_f32<TIn>(tb[static_cast<size_t>(p) * C + c]) : 0.f;
  }
  __syncthreads();
  TOut* ob = out + static_cast<size_t>(b) * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) ob[static_cast<size_t>(c) * HW + p] = from_f32<TOut>(tile[threadIdx.x][i]);
  }
}

// Fast path for HW % 8 == 0 and C % 8 == 0: 64 x 64 tiles, 16-byte global accesses on both sides.
// tokens [B*HW, C] bf16 -> NCHW bf16 [B,C,HW].
static __global__ void __launch_bounds__(256)
unpack_tokens_bf16_fast_kernel(const bf16* __restrict__ tok, bf16* __restrict__ out, int C, int HW) {
  __shared__ uint16_t tile[64][66];  // [hw][c]
  const int b = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const uint16_t* tb = reinterpret_cast<const uint16_t*>(tok) + static_cast<size_t>(b) * HW * C;
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {
    const int p = i >> 3, co = (i & 7) * 8;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (p0 + p < HW && c0 + co < C) q = __ldg(reinterpret_cast<const uint4*>(tb + static_cast<size_t>(p0 + p) * C + c0 + co));
    uint32_t* dst = reinterpret_cast<uint32_t*>(&tile[p][co]);
    dst[0] = q.x; dst[1] = q.y; dst[2] = q.z; dst[3] = q.w;
  }
  __syncthreads();
  uint16_t* ob = reinterpret_cast<uint16_t*>(out) + static_cast<size_t>(b) * C * HW;
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {
    const int c = i >> 3, po = (i & 7) * 8;
    if (c0 + c < C && p0 + po < HW) {
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        w[k] = static_cast<uint32_t>(tile[po + 2 * k][c]) | (static_cast<uint32_t>(tile[po + 2 * k + 1][c]) << 16);
      *reinterpret_cast<uint4*>(ob + static_cast<size_t>(c0 + c) * HW + p0 + po) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// ---- 64 x 64 tile movers between NCHW and token-major ------------------------------------------------------------
// 8 consecutive spatial positions of one channel row, read / written with the widest access the row pitch allows:
// VEC = 8 (HW % 8 == 0), 4 (HW % 4 == 0) or 1 elements per access.  nv = number of valid positions (0..8).
template <typename T, int VEC> struct Row8;
template <int VEC> struct Row8<bf16, VEC> {
  static __device__ __forceinline__ void load(const bf16* p, int nv, float (&o)[8]) {
    const uint16_t* q = reinterpret_cast<const uint16_t*>(p);
    if (VEC == 8) {
      uint4 w = make_uint4(0, 0, 0, 0);
      if (nv > 0) w = __ldg(reinterpret_cast<const uint4*>(q));
      o[0] = bf16lo(w.x); o[1] = bf16hi(w.x); o[2] = bf16lo(w.y); o[3] = bf16hi(w.y);
      o[4] = bf16lo(w.z); o[5] = bf16hi(w.z); o[6] = bf16lo(w.w); o[7] = bf16hi(w.w);
    } else if (VEC == 4) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint2 w = make_uint2(0, 0);
        if (nv > 4 * h) w = __ldg(reinterpret_cast<const uint2*>(q + 4 * h));
        o[4 * h] = bf16lo(w.x); o[4 * h + 1] = bf16hi(w.x); o[4 * h + 2] = bf16lo(w.y); o[4 * h + 3] = bf16hi(w.y);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = j < nv ? __uint_as_float(static_cast<uint32_t>(__ldg(q + j)) << 16) : 0.f;
    }
  }
  static __device__ __forceinline__ void store(bf16* p, int nv, const float (&v)[8]) {
    uint16_t* q = reinterpret_cast<uint16_t*>(p);
    if (VEC == 8) {
      if (nv > 0)
        *reinterpret_cast<uint4*>(q) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                  pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    } else if (VEC == 4) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (nv > 4 * h)
          *reinterpret_cast<uint2*>(q + 4 * h) = make_uint2(pack_bf16x2(v[4 * h], v[4 * h + 1]), pack_bf16x2(v[4 * h + 2], v[4 * h + 3]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nv) p[j] = __float2bfloat16_rn(v[j]);
    }
  }
};
template <int VEC> struct Row8<float, VEC> {
  static __device__ __forceinline__ void load(const float* p, int nv, float (&o)[8]) {
    if (VEC >= 4) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nv > 4 * h) w = __ldg(reinterpret_cast<const float4*>(p + 4 * h));
        o[4 * h] = w.x; o[4 * h + 1] = w.y; o[4 * h + 2] = w.z; o[4 * h + 3] = w.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = j < nv ? __ldg(p + j) : 0.f;
    }
  }
  static __device__ __forceinline__ void store(float* p, int nv, const float (&v)[8]) {
    if (VEC >= 4) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (nv > 4 * h) *reinterpret_cast<float4*>(p + 4 * h) = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nv) p[j] = v[j];
    }
  }
};

// NCHW [B,C,HW] (bf16 or fp32) -> bf16 tokens [B*HW, C]; sae_mlp.py:44 'b c h w -> (b h w) c'.
// grid (ceil(HW/64), ceil(C/64), B), 256 threads.  A thread reads 8 positions of TWO neighbouring channels, interleaves
// them into 8 words (channel pair of one position each) and stores those into a [64 positions][32 channel pairs] tile
// whose 16-byte groups are XOR-swizzled by the position octet, so that both the 4-byte writes and the 16-byte reads of
// the second phase (8 channels of one position = one output store) are bank-conflict free.  (The first version
// gathered 8 two-byte values per output store and was issue- / LSU-bound at 52 us; profiles/r01g.)
// slab_rows > 0: write the slab-major layout [ceil(C/64)][slab_rows = B*HW][64] (gemm_host.cuh) instead of [B*HW, C];
// the padding columns of the last slab (C % 64 != 0) are written as zeros.
// xpart != null: also emit, per image, HW tile and channel, the statistics of the bf16-rounded x that the loss
// metrics need (sum, sum of squares, min, max): xpart[((b * gridDim.x + tile) * 4 + q) * C + c].
template <typename TIn, int VEC>
static __global__ void __launch_bounds__(256)
pack_nchw_tile_kernel(const TIn* __restrict__ x, bf16* __restrict__ out, int C, int HW, long long slab_rows,
                      float* __restrict__ xpart) {
  __shared__ __align__(16) uint32_t tile[64][32];  // [position][channel pair], groups of 4 words swizzled
  const int b = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 64;
  const TIn* xb = x + static_cast<size_t>(b) * C * HW;
  {
    const int c2 = threadIdx.x >> 3, oct = threadIdx.x & 7, po = oct * 8;
    const int ch = c0 + 2 * c2;                      // C is even: ch < C <=> ch + 1 < C
    const int nv = ch < C ? max(0, min(8, HW - (p0 + po))) : 0;
    uint32_t w[8];                                   // w[k] = (x[ch][po + k], x[ch + 1][po + k]) as bf16 pairs
    if (sizeof(TIn) == 2 && VEC == 8) {              // bf16, 16-byte aligned rows: no conversion at all
      uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
      if (nv > 0) {
        q0 = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(ch) * HW + p0 + po));
        q1 = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(ch + 1) * HW + p0 + po));
      }
      w[0] = __byte_perm(q0.x, q1.x, 0x5410); w[1] = __byte_perm(q0.x, q1.x, 0x7632);
      w[2] = __byte_perm(q0.y, q1.y, 0x5410); w[3] = __byte_perm(q0.y, q1.y, 0x7632);
      w[4] = __byte_perm(q0.z, q1.z, 0x5410); w[5] = __byte_perm(q0.z, q1.z, 0x7632);
      w[6] = __byte_perm(q0.w, q1.w, 0x5410); w[7] = __byte_perm(q0.w, q1.w, 0x7632);
    } else {
      float v0[8], v1[8];
      Row8<TIn, VEC>::load(xb + static_cast<size_t>(ch) * HW + p0 + po, nv, v0);
      Row8<TIn, VEC>::load(xb + static_cast<size_t>(ch + 1) * HW + p0 + po, nv, v1);
#pragma unroll
      for (int k = 0; k < 8; ++k) w[k] = pack_bf16x2(v0[k], v1[k]);
    }
    const int col = (((c2 >> 2) ^ oct) << 2) | (c2 & 3);   // (po + k) >> 3 == oct for every k
#pragma unroll
    for (int k = 0; k < 8; ++k) tile[po + k][col] = w[k];
    if (xpart) {  // the 8 lanes that share the channel pair combine their 8 positions each (fixed butterfly order)
      float s0 = 0.f, q0 = 0.f, mn0 = INFINITY, mx0 = -INFINITY, s1 = 0.f, q1 = 0.f, mn1 = INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k < nv) {
          const float lo = bf16lo(w[k]), hi = bf16hi(w[k]);
          s0 += lo; q0 += lo * lo; mn0 = fminf(mn0, lo); mx0 = fmaxf(mx0, lo);
          s1 += hi; q1 += hi * hi; mn1 = fminf(mn1, hi); mx1 = fmaxf(mx1, hi);
        }
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o); q0 += __shfl_xor_sync(0xffffffffu, q0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o); q1 += __shfl_xor_sync(0xffffffffu, q1, o);
        mn0 = fminf(mn0, __shfl_xor_sync(0xffffffffu, mn0, o)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
        mn1 = fminf(mn1, __shfl_xor_sync(0xffffffffu, mn1, o)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
      }
      if (oct == 0 && ch < C) {
        float* o = xpart + (static_cast<size_t>(b) * gridDim.x + blockIdx.x) * 4 * C + ch;
        *reinterpret_cast<float2*>(o) = make_float2(s0, s1);
        *reinterpret_cast<float2*>(o + C) = make_float2(q0, q1);
        *reinterpret_cast<float2*>(o + 2 * C) = make_float2(mn0, mn1);
        *reinterpret_cast<float2*>(o + 3 * C) = make_float2(mx0, mx1);
      }
    }
  }
  __syncthreads();
  uint16_t* ob = reinterpret_cast<uint16_t*>(out);
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int j = threadIdx.x + 256 * it;
    const int p = j >> 3, g = j & 7, co = g * 8;    // 8 channels of one position: one 16-byte group of the tile row
    if (p0 + p < HW && (c0 + co < C || slab_rows > 0)) {   // tile rows of channels >= C hold zeros
      const uint4 q = *reinterpret_cast<const uint4*>(&tile[p][(g ^ (p >> 3)) << 2]);
      const size_t t = static_cast<size_t>(b) * HW + p0 + p;
      const size_t off = slab_rows > 0 ? (static_cast<size_t>(blockIdx.y) * slab_rows + t) * 64 + co : t * C + c0 + co;
      *reinterpret_cast<uint4*>(ob + off) = q;
    }
  }
}

// Token-major activations X [B*HW, C] (bf16; what a channels_last base model emits, read zero-copy by the GEMMs) ->
// the statistics the loss metrics need, in the pack kernel's format: xpart[((b * NT + tile) * 4 + q) * C + c] with
// q = sum x, sum x^2, min x, max x over the 64 positions of HW tile `tile` of image b.  One WARP per (image, tile,
// 256-channel group): lane owns 8 consecutive channels (one 16-byte load per row, 8 rows in flight), walks the 64 rows
// in order and writes its results straight from registers -- no shared memory, no synchronisation, read-only and
// HBM-bound (103 MB at cfg2).  grid = ceil(B * NT * ceil(C/256) / 4) blocks of 4 warps.
static __global__ void __launch_bounds__(128)
x_stats_tokens_kernel(const bf16* __restrict__ x, float* __restrict__ xpart, int C, int HW, int NT, long long n_warps) {
  const long long wid = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (wid >= n_warps) return;
  const int lane = threadIdx.x & 31;
  const int cgroups = (C + 255) / 256;
  const int cg = static_cast<int>(wid % cgroups);
  const long long bt = wid / cgroups;                 // b * NT + tile
  const int tile = static_cast<int>(bt % NT);
  const long long b = bt / NT;
  const int c0 = cg * 256 + lane * 8;
  if (c0 >= C) return;
  const int p0 = tile * 64, np = min(64, HW - p0);
  const bf16* xb = x + (static_cast<size_t>(b) * HW + p0) * C + c0;
  float s[8], q[8], mn[8], mx[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s[k] = 0.f; q[k] = 0.f; mn[k] = INFINITY; mx[k] = -INFINITY; }
  for (int r0 = 0; r0 < np; r0 += 8) {
    uint4 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] = r0 + i < np ? __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(r0 + i) * C)) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (r0 + i < np) {
        const uint32_t ws[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float lo = bf16lo(ws[k]), hi = bf16hi(ws[k]);
          s[2 * k] += lo; q[2 * k] += lo * lo; mn[2 * k] = fminf(mn[2 * k], lo); mx[2 * k] = fmaxf(mx[2 * k], lo);
          s[2 * k + 1] += hi; q[2 * k + 1] += hi * hi; mn[2 * k + 1] = fminf(mn[2 * k + 1], hi); mx[2 * k + 1] = fmaxf(mx[2 * k + 1], hi);
        }
      }
    }
  }
  float* o = xpart + static_cast<size_t>(bt) * 4 * C + c0;
  *reinterpret_cast<float4*>(o) = make_float4(s[0], s[1], s[2], s[3]);
  *reinterpret_cast<float4*>(o + 4) = make_float4(s[4], s[5], s[6], s[7]);
  *reinterpret_cast<float4*>(o + C) = make_float4(q[0], q[1], q[2], q[3]);
  *reinterpret_cast<float4*>(o + C + 4) = make_float4(q[4], q[5], q[6], q[7]);
  *reinterpret_cast<float4*>(o + 2 * C) = make_float4(mn[0], mn[1], mn[2], mn[3]);
  *reinterpret_cast<float4*>(o + 2 * C + 4) = make_float4(mn[4], mn[5], mn[6], mn[7]);
  *reinterpret_cast<float4*>(o + 3 * C) = make_float4(mx[0], mx[1], mx[2], mx[3]);
  *reinterpret_cast<float4*>(o + 3 * C + 4) = make_float4(mx[4], mx[5], mx[6], mx[7]);
}
inline void launch_x_stats_tokens(cudaStream_t st, const bf16* x, float* xpart, int C, int HW, int NT, long long n_img) {
  const long long n_warps = n_img * NT * ((C + 255) / 256);
  (x_stats_tokens_kernel<<<static_cast<unsigned>((n_warps + 3) / 4), 128, 0, st>>>(x, xpart, C, HW, NT, n_warps), svb::count_launch());
}

// Channel-major decoder output [C][ld] (tokens t = b*HW + p contiguous per channel) -> the caller's NCHW tensor
// [B, C, HW]: every image's HW-long run of a channel is copied as it is.  grid (ceil(T / (1024 * VEC)), C), 256 threads,
// 4 pieces of VEC elements per thread; VEC = 4 needs HW % 4 == 0 (a piece never straddles two images) and an 8-byte
// (bf16) / 16-byte (fp32) aligned output.
template <typename TOut, int VEC>
static __global__ void __launch_bounds__(256)
cmajor_to_nchw_kernel(const bf16* __restrict__ dt, TOut* __restrict__ out, int C, int HW, long long T, long long ld) {
  const int c = blockIdx.y;
  const uint16_t* src = reinterpret_cast<const uint16_t*>(dt) + static_cast<size_t>(c) * ld;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long t = (blockIdx.x * 1024LL + k * 256 + threadIdx.x) * VEC;
    if (t < T) {
      const long long b = t / HW;
      const int p = static_cast<int>(t - b * HW);
      const size_t o = (static_cast<size_t>(b) * C + c) * HW + p;
      if (VEC == 4) {
        const uint2 w = *reinterpret_cast<const uint2*>(src + t);
        if (sizeof(TOut) == 2) *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(out) + o) = w;
        else *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = make_float4(bf16lo(w.x), bf16hi(w.x), bf16lo(w.y), bf16hi(w.y));
      } else {
        const uint16_t h = src[t];
        if (sizeof(TOut) == 2) reinterpret_cast<uint16_t*>(out)[o] = h;
        else reinterpret_cast<float*>(out)[o] = __uint_as_float(static_cast<uint32_t>(h) << 16);
      }
    }
  }
}

// Fused "after the decoder" pass for NCHW inputs: ONE sweep over d (token-major bf16) and x (NCHW) that
//   * writes d back in NCHW (what the hook returns, model_pipeline.py:425,432; utils.py:2478) when d_out != null, and
//   * accumulates, per image and channel, sum x, sum x^2, sum d, sum d^2, sum (d-x), sum (d-x)^2, min x, max x
//     (variance_explained utils.py:2012-2030, compute_rmse_nrmse sparse_loss.py:4-21, db_dec).
// x is rounded to bf16 first, like the GEMM operand.  st layout: [b][chunk][8][C] (the channel_stats layout).
// grid (ceil(C/64), B, chunks), 256 threads: a block walks tiles_per_chunk HW tiles of its 64 channels; thread i owns
// channel (i>>3) [+32] and 8 consecutive positions per tile, so the statistics stay in registers until the end.
template <typename TIn, typename TOut, int VEC>
static __global__ void __launch_bounds__(256, 3)
post_dec_nchw_kernel(const bf16* __restrict__ d_tok, const TIn* __restrict__ x, TOut* __restrict__ d_out,
                     float* __restrict__ st, int C, int HW, int tiles_per_chunk, long long slab_rows) {
  __shared__ uint16_t tile[64][66];  // [hw][c]
  const int b = blockIdx.y, c0 = blockIdx.x * 64, rc = blockIdx.z, R = gridDim.z;
  // d_tok: [B*HW, C] or, slab_rows > 0, slab-major [C/64][slab_rows][64]; tb points at (token b*HW, channel c0)
  const size_t d_pitch = slab_rows > 0 ? 64 : static_cast<size_t>(C);
  const uint16_t* tb = reinterpret_cast<const uint16_t*>(d_tok) +
                       (slab_rows > 0 ? (static_cast<size_t>(blockIdx.x) * slab_rows + static_cast<size_t>(b) * HW) * 64
                                      : static_cast<size_t>(b) * HW * C + c0);
  const TIn* xb = x + static_cast<size_t>(b) * C * HW;
  TOut* ob = d_out ? d_out + static_cast<size_t>(b) * C * HW : nullptr;
  const int p_begin = rc * tiles_per_chunk * 64, p_end = min(HW, p_begin + tiles_per_chunk * 64);
  float a[2][8];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
#pragma unroll
    for (int q = 0; q < 6; ++q) a[k][q] = 0.f;
    a[k][6] = INFINITY;
    a[k][7] = -INFINITY;
  }
  // Both streams of a tile (16 bytes of d and 8 positions of x per thread and half) are fetched one tile ahead.
  uint4 dq[2];
  float xv[2][8];
  int nv[2];
  auto fetch = [&](int p0) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = threadIdx.x + 256 * k;
      const int p = i >> 3, co = (i & 7) * 8;
      dq[k] = make_uint4(0, 0, 0, 0);
      if (p0 + p < p_end && c0 + co < C) dq[k] = __ldg(reinterpret_cast<const uint4*>(tb + static_cast<size_t>(p0 + p) * d_pitch + co));
      const int c = i >> 3, po = (i & 7) * 8;
      nv[k] = (c0 + c < C) ? max(0, min(8, p_end - (p0 + po))) : 0;
      Row8<TIn, VEC>::load(xb + static_cast<size_t>(c0 + c) * HW + p0 + po, nv[k], xv[k]);
    }
  };
  if (p_begin < p_end) fetch(p_begin);
  for (int p0 = p_begin; p0 < p_end; p0 += 64) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = threadIdx.x + 256 * k;
      uint32_t* dst = reinterpret_cast<uint32_t*>(&tile[i >> 3][(i & 7) * 8]);
      dst[0] = dq[k].x; dst[1] = dq[k].y; dst[2] = dq[k].z; dst[3] = dq[k].w;
    }
    float xc[2][8];
    int nc[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      nc[k] = nv[k];
#pragma unroll
      for (int j = 0; j < 8; ++j) xc[k][j] = xv[k][j];
    }
    if (p0 + 64 < p_end) fetch(p0 + 64);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = threadIdx.x + 256 * k;
      const int c = i >> 3, po = (i & 7) * 8;
      if (nc[k] == 0) continue;
      float dv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dv[j] = __uint_as_float(static_cast<uint32_t>(tile[po + j][c]) << 16);
      if (ob) Row8<TOut, VEC>::store(ob + static_cast<size_t>(c0 + c) * HW + p0 + po, nc[k], dv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < nc[k]) {
          const float xr = sizeof(TIn) == 2 ? xc[k][j] : __bfloat162float(__float2bfloat16_rn(xc[k][j]));
          const float f = dv[j] - xr;
          a[k][0] += xr; a[k][1] += xr * xr;
          a[k][2] += dv[j]; a[k][3] += dv[j] * dv[j];
          a[k][4] += f; a[k][5] += f * f;
          a[k][6] = fminf(a[k][6], xr); a[k][7] = fmaxf(a[k][7], xr);
        }
      }
    }
  }
  // the 8 lanes that share a channel combine in a fixed (butterfly) order
#pragma unroll
  for (int k = 0; k < 2; ++k) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float v = a[k][q];
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const float u = __shfl_xor_sync(0xffffffffu, v, o);
        v = q < 6 ? v + u : (q == 6 ? fminf(v, u) : fmaxf(v, u));
      }
      const int c = c0 + (threadIdx.x >> 3) + 32 * k;
      if ((threadIdx.x & 7) == 0 && c < C) st[((static_cast<size_t>(b) * R + rc) * 8 + q) * C + c] = v;
    }
  }
}

template <typename TIn, typename TOut>
static __global__ void convert_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = from_f32<TOut>(to_f32<TIn>(in[i]));
}

static __global__ void fill_u32_kernel(uint32_t* p, size_t n, uint32_t v) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}

// ------------------------------------------------------------------------------------------------ weights
// One warp per encoder row f:  w_bf16[f,:] = bf16(w[f,:]);  fold[f] = b_enc[f] - sum_c bf16(w[f,c]) * b_dec[c]
// (the pre-bias subtraction x - b_dec of sae_mlp.py:49 folded into the encoder bias).  dotw (optional) returns
// the raw dot product sum_c bf16(w[f,c]) * b_dec[c] (used by the gated path).
static __global__ void prep_encoder_kernel(const float* __restrict__ w, const float* __restrict__ b_enc,
                                    const float* __restrict__ b_dec, bf16* __restrict__ w_bf16,
                                    float* __restrict__ fold, float* __restrict__ dotw, int F, int C) {
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (f >= F) return;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) {
    const bf16 q = __float2bfloat16_rn(w[static_cast<size_t>(f) * C + c]);
    w_bf16[static_cast<size_t>(f) * C + c] = q;
    acc += __bfloat162float(q) * b_dec[c];
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if (fold) fold[f] = (b_enc ? b_enc[f] : 0.f) - acc;
    if (dotw) dotw[f] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ reductions
// out[j] = scale * sum_{i<R} in[i*ld + j], rows summed in a fixed order.  Two stages when R is large:
// stage 1 (gridDim.y row chunks) writes [gridDim.y, N] partials, stage 2 finishes them.
static __global__ void reduce_rows_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int N, size_t ld,
                                   float scale) {
  __shared__ float s[8][33];
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const int g = threadIdx.x >> 5;  // 8 row lanes
  const int chunks = gridDim.y;
  const int rows_per = (R + chunks - 1) / chunks;
  const int r0 = blockIdx.y * rows_per, r1 = min(R, r0 + rows_per);
  float acc = 0.f;
  if (j < N)
    for (int i = r0 + g; i < r1; i += 8) acc += in[static_cast<size_t>(i) * ld + j];
  s[g][threadIdx.x & 31] = acc;
  __syncthreads();
  if (g == 0 && j < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s[k][threadIdx.x & 31];
    out[static_cast<size_t>(blockIdx.y) * N + j] = t * scale;
  }
}

// Sum of a flat array in a fixed order, one block: out[0] = scale * sum(in[0..n)).
static __global__ void reduce_flat_kernel(const float* __restrict__ in, size_t n, float scale, float* __restrict__ out) {
  __shared__ float s[32];
  float acc = 0.f;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) acc += in[i];
  const float r = block_sum(acc, s);
  if (threadIdx.x == 0) out[0] = r * scale;
}

// ------------------------------------------------------------------------------------------------ channel stats
// Per image b, row chunk r and channel c over the chunk's tokens (token-major bf16 inputs):
//   st[b][r][0][c] = sum x, [1] = sum x^2, [2] = sum d, [3] = sum d^2, [4] = sum diff, [5] = sum diff^2,
//   [6] = min x, [7] = max x          (diff = d - x, from the stored bf16 d)
// grid (n_img, row_chunks, ceil(C/256)), 256 threads: lane owns 8 channels (one 16-byte load), warp w takes rows
// w, w+8, ... of the chunk.  Feeds variance_explained (utils.py:2012-2030), compute_rmse_nrmse
// (sparse_loss.py:4-21) and db_dec.
static __global__ void __launch_bounds__(256)
channel_stats_kernel(const bf16* __restrict__ x, const bf16* __restrict__ d, float* __restrict__ st, int C, int HW,
                     int rows_per_chunk) {
  __shared__ float s[8][8][33];  // [warp][stat][lane]  (one channel-of-8 at a time)
  const int b = blockIdx.x, rc = blockIdx.y, R = gridDim.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c0 = blockIdx.z * 256 + lane * 8;
  const int r_begin = rc * rows_per_chunk, r_end = min(HW, r_begin + rows_per_chunk);
  float a[8][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int q = 0; q < 6; ++q) a[q][k] = 0.f;
    a[6][k] = INFINITY;
    a[7][k] = -INFINITY;
  }
  auto accumulate = [&](const uint4& qx, const uint4& qd) {
    const uint32_t wx[4] = {qx.x, qx.y, qx.z, qx.w}, wd[4] = {qd.x, qd.y, qd.z, qd.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float x0 = bf16lo(wx[k]), x1 = bf16hi(wx[k]);
      const float d0 = bf16lo(wd[k]), d1 = bf16hi(wd[k]);
      const float f0 = d0 - x0, f1 = d1 - x1;
      a[0][2 * k] += x0; a[0][2 * k + 1] += x1;
      a[1][2 * k] += x0 * x0; a[1][2 * k + 1] += x1 * x1;
      a[2][2 * k] += d0; a[2][2 * k + 1] += d1;
      a[3][2 * k] += d0 * d0; a[3][2 * k + 1] += d1 * d1;
      a[4][2 * k] += f0; a[4][2 * k + 1] += f1;
      a[5][2 * k] += f0 * f0; a[5][2 * k + 1] += f1 * f1;
      a[6][2 * k] = fminf(a[6][2 * k], x0); a[6][2 * k + 1] = fminf(a[6][2 * k + 1], x1);
      a[7][2 * k] = fmaxf(a[7][2 * k], x0); a[7][2 * k + 1] = fmaxf(a[7][2 * k + 1], x1);
    }
  };
  if (c0 < C) {
    const bf16* xb = x + static_cast<size_t>(b) * HW * C + c0;
    const bf16* db = d + static_cast<size_t>(b) * HW * C + c0;
    int r = r_begin + w;
    for (; r + 24 < r_end; r += 32) {  // 4 rows (8 independent 16-byte loads) in flight per thread
      uint4 qx[4], qd[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        qx[u] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(r + 8 * u) * C));
        qd[u] = __ldg(reinterpret_cast<const uint4*>(db + static_cast<size_t>(r + 8 * u) * C));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) accumulate(qx[u], qd[u]);
    }
    for (; r < r_end; r += 8)
      accumulate(__ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(r) * C)),
                 __ldg(reinterpret_cast<const uint4*>(db + static_cast<size_t>(r) * C)));
  }
  // cross-warp combine, one of the 8 per-lane channels at a time (fixed order over warps)
  for (int k = 0; k < 8; ++k) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) s[w][q][lane] = a[q][k];
    __syncthreads();
    if (w == 0 && c0 + k < C) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float t = s[0][q][lane];
        for (int ww = 1; ww < 8; ++ww) {
          const float o = s[ww][q][lane];
          t = q < 6 ? t + o : (q == 6 ? fminf(t, o) : fmaxf(t, o));
        }
        st[((static_cast<size_t>(b) * R + rc) * 8 + q) * C + c0 + k] = t;
      }
    }
  }
}

// Per-image channel statistics when the decoder epilogue (EpiDecNchw) and the pack kernel produced the partials:
//   dpart[((g * 2 + slot) * 3 + q) * C + c]   g = 32-token row group, slot = image index relative to the group's first
//                                             image, q = sum d, sum d^2 (per image) and sum (d-x)^2 (whole group, slot 0)
//   xpart[((b * NT + tile) * 4 + q) * C + c]  q = sum x, sum x^2, min x, max x per HW tile of 64 positions
// Stage 1, grid (B, ceil(C/64)), 256 threads = 4 group lanes x 64 channels (a warp reads 32 consecutive channels):
//   vb[(b * 6 + q) * C + c] = Var_hw(x), Var_hw(d) (unbiased, utils.py:2015,2020), sum (d-x), sum (d-x)^2, min x, max x.
// A group's sum (d-x)^2 is credited to its first image; only its total over images is used downstream.
static __global__ void __launch_bounds__(256)
dec_stats_image_kernel(const float* __restrict__ dpart, const float* __restrict__ xpart, float* __restrict__ vb, int C,
                       int HW, int NT, long long T) {
  __shared__ float sh[4][7][64];
  const int b = blockIdx.x, cl = threadIdx.x & 63, gl = threadIdx.x >> 6;
  const int c = blockIdx.y * 64 + cl;
  const long long t0 = static_cast<long long>(b) * HW, t1 = min(t0 + HW, T) - 1;
  const long long g0 = t0 >> 5, g1 = t1 >> 5;
  float sd = 0.f, sd2 = 0.f, sq = 0.f, sx = 0.f, sx2 = 0.f, mn = INFINITY, mx = -INFINITY;
  if (c < C) {
    for (long long g = g0 + gl; g <= g1; g += 4) {
      const int slot = b - static_cast<int>((g << 5) / HW);
      const float* p = dpart + (static_cast<size_t>(g) * 2 + slot) * 3 * C + c;
      sd += p[0];
      sd2 += p[C];
      if (slot == 0) sq += p[2 * C];
    }
    for (int t = gl; t < NT; t += 4) {
      const float* p = xpart + (static_cast<size_t>(b) * NT + t) * 4 * C + c;
      sx += p[0];
      sx2 += p[C];
      mn = fminf(mn, p[2 * C]);
      mx = fmaxf(mx, p[3 * C]);
    }
  }
  sh[gl][0][cl] = sd; sh[gl][1][cl] = sd2; sh[gl][2][cl] = sq; sh[gl][3][cl] = sx; sh[gl][4][cl] = sx2;
  sh[gl][5][cl] = mn; sh[gl][6][cl] = mx;
  __syncthreads();
  if (gl == 0 && c < C) {
    for (int k = 1; k < 4; ++k) {
      sd += sh[k][0][cl]; sd2 += sh[k][1][cl]; sq += sh[k][2][cl]; sx += sh[k][3][cl]; sx2 += sh[k][4][cl];
      mn = fminf(mn, sh[k][5][cl]); mx = fmaxf(mx, sh[k][6][cl]);
    }
    const float inv = 1.f / static_cast<float>(HW), invm1 = HW > 1 ? 1.f / static_cast<float>(HW - 1) : 0.f;
    float* o = vb + static_cast<size_t>(b) * 6 * C + c;
    o[0] = fmaxf(sx2 - sx * sx * inv, 0.f) * invm1;
    o[C] = fmaxf(sd2 - sd * sd * inv, 0.f) * invm1;
    o[2 * C] = sd - sx;
    o[3 * C] = sq;
    o[4 * C] = mn;
    o[5 * C] = mx;
  }
}
// Stage 2, grid ceil(C/32), 1024 threads = 32 image lanes x 32 channels: the same outputs as
// channel_stats_finalize_kernel (chan[4][C] and one (sum Var x, sum Var d) pair per block), images in a fixed order.
static __global__ void __launch_bounds__(1024)
dec_stats_channel_kernel(const float* __restrict__ vb, float* __restrict__ chan, float* __restrict__ var_partial, int B,
                         int C) {
  __shared__ float sh[32][6][33];
  __shared__ float sv[2][32];
  const int cl = threadIdx.x & 31, bl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float vx = 0.f, vd = 0.f, sdf = 0.f, sq = 0.f, mn = INFINITY, mx = -INFINITY;
  if (c < C)
    for (int b = bl; b < B; b += 32) {
      const float* p = vb + static_cast<size_t>(b) * 6 * C + c;
      vx += p[0]; vd += p[C]; sdf += p[2 * C]; sq += p[3 * C];
      mn = fminf(mn, p[4 * C]); mx = fmaxf(mx, p[5 * C]);
    }
  sh[bl][0][cl] = vx; sh[bl][1][cl] = vd; sh[bl][2][cl] = sdf; sh[bl][3][cl] = sq; sh[bl][4][cl] = mn; sh[bl][5][cl] = mx;
  __syncthreads();
  if (bl < 6) {  // warp q combines statistic q over the 32 image lanes
    const int q = bl;
    float t = sh[0][q][cl];
    for (int k = 1; k < 32; ++k) {
      const float o = sh[k][q][cl];
      t = q < 4 ? t + o : (q == 4 ? fminf(t, o) : fmaxf(t, o));
    }
    if (q >= 2) {
      if (c < C) chan[(q - 2) * C + c] = t;
    } else {
      sv[q][cl] = c < C ? t : 0.f;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, bb = 0.f;
    for (int k = 0; k < 32; ++k) { a += sv[0][k]; bb += sv[1][k]; }
    var_partial[blockIdx.x * 2] = a;
    var_partial[blockIdx.x * 2 + 1] = bb;
  }
}

// Collapse [B][R][8][C] chunk stats into: chan[0][c] = sum diff, chan[1][c] = sum diff^2, chan[2][c] = min x,
// chan[3][c] = max x;  var_partial[blk] = (sum_{b,c in blk} Var_hw(x), Var_hw(d))  (unbiased, utils.py:2015,2020).
// One block per 8 channels: 8 channel lanes x 32 image lanes; chunks and images combined in a fixed order.
static __global__ void __launch_bounds__(256)
channel_stats_finalize_kernel(const float* __restrict__ st, float* __restrict__ chan,
                              float* __restrict__ var_partial /* [gridDim.x][2] */, int B, int R, int C, int HW) {
  __shared__ float s[6][8][33];
  const int cl = threadIdx.x & 7, bl = threadIdx.x >> 3;  // channel lane, image lane
  const int c = blockIdx.x * 8 + cl;
  float sd = 0.f, sd2 = 0.f, mn = INFINITY, mx = -INFINITY, vx = 0.f, vd = 0.f;
  if (c < C) {
    const float inv = 1.f / static_cast<float>(HW), invm1 = HW > 1 ? 1.f / static_cast<float>(HW - 1) : 0.f;
    for (int b = bl; b < B; b += 32) {
      float sx = 0.f, sx2 = 0.f, sdd = 0.f, sdd2 = 0.f;
      for (int r = 0; r < R; ++r) {
        const float* p = st + (static_cast<size_t>(b) * R + r) * 8 * C + c;
        sx += p[0]; sx2 += p[C]; sdd += p[2 * C]; sdd2 += p[3 * C];
        sd += p[4 * C];
        sd2 += p[5 * C];
        mn = fminf(mn, p[6 * C]);
        mx = fmaxf(mx, p[7 * C]);
      }
      vx += fmaxf(sx2 - sx * sx * inv, 0.f) * invm1;
      vd += fmaxf(sdd2 - sdd * sdd * inv, 0.f) * invm1;
    }
  }
  s[0][cl][bl] = sd; s[1][cl][bl] = sd2; s[2][cl][bl] = mn; s[3][cl][bl] = mx; s[4][cl][bl] = vx; s[5][cl][bl] = vd;
  __syncthreads();
  if (bl == 0) {
    float t[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) t[q] = s[q][cl][0];
    for (int k = 1; k < 32; ++k) {
      t[0] += s[0][cl][k]; t[1] += s[1][cl][k];
      t[2] = fminf(t[2], s[2][cl][k]); t[3] = fmaxf(t[3], s[3][cl][k]);
      t[4] += s[4][cl][k]; t[5] += s[5][cl][k];
    }
    if (c < C) {
      chan[c] = t[0]; chan[C + c] = t[1]; chan[2 * C + c] = t[2]; chan[3 * C + c] = t[3];
    } else {
      t[4] = 0.f; t[5] = 0.f;
    }
    // threads 0..7 (bl == 0) hold the block's 8 channels: fixed-order sum
    s[4][cl][0] = t[4];
    s[5][cl][0] = t[5];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, bb = 0.f;
    for (int k = 0; k < 8; ++k) { a += s[4][k][0]; bb += s[5][k][0]; }
    var_partial[blockIdx.x * 2] = a;
    var_partial[blockIdx.x * 2 + 1] = bb;
  }
}

// 2-D inputs (hw == 1): variance_explained takes the variance over the feature axis of each row
// (utils.py:2022-2027).  One warp per row; rowvar[r*2] = Var_c(x[r,:]), [r*2+1] = Var_c(d[r,:]).
static __global__ void row_variance_kernel(const bf16* __restrict__ x, const bf16* __restrict__ d, float* __restrict__ rowvar,
                                    int T, int C) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= T) return;
  float sx = 0.f, sx2 = 0.f, sd = 0.f, sd2 = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float a = __bfloat162float(x[static_cast<size_t>(r) * C + c]);
    const float b = __bfloat162float(d[static_cast<size_t>(r) * C + c]);
    sx += a; sx2 += a * a; sd += b; sd2 += b * b;
  }
  sx = warp_sum(sx); sx2 = warp_sum(sx2); sd = warp_sum(sd); sd2 = warp_sum(sd2);
  if (lane == 0) {
    const float inv = 1.f / C, invm1 = C > 1 ? 1.f / (C - 1) : 0.f;
    rowvar[2 * r] = fmaxf(sx2 - sx * sx * inv, 0.f) * invm1;
    rowvar[2 * r + 1] = fmaxf(sd2 - sd * sd * inv, 0.f) * invm1;
  }
}

// ------------------------------------------------------------------------------------------------ activity
// Per-image activity bits (utils.py:2033-2047: a unit is active for an image iff any of its pixels is non-zero) from
// the encoder's group-major 1-bit masks (epilogues.cuh: mask_index): bits[b][4*wg + i] = OR over the image's rows of
// word i of group wg.  One coalesced pass over the 26 MB mask on the side stream replaces ~400 k global atomics and
// the warp-wide ORs in the encoder epilogue (which cost that GEMM 0.025 ms).  grid (n_img, word groups), 128 threads.
static __global__ void __launch_bounds__(128)
mask_to_activity_kernel(const uint32_t* __restrict__ mask, uint32_t* __restrict__ bits, long long T, int hw, int words) {
  __shared__ uint32_t sh[4][4];
  const int b = blockIdx.x, wg = blockIdx.y;
  const uint4* src = reinterpret_cast<const uint4*>(mask) + static_cast<size_t>(wg) * T + static_cast<size_t>(b) * hw;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int r = threadIdx.x; r < hw; r += 128) {
    const uint4 v = __ldg(src + r);
    acc.x |= v.x; acc.y |= v.y; acc.z |= v.z; acc.w |= v.w;
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  acc.x = __reduce_or_sync(0xffffffffu, acc.x); acc.y = __reduce_or_sync(0xffffffffu, acc.y);
  acc.z = __reduce_or_sync(0xffffffffu, acc.z); acc.w = __reduce_or_sync(0xffffffffu, acc.w);
  if (lane == 0) { sh[w][0] = acc.x; sh[w][1] = acc.y; sh[w][2] = acc.z; sh[w][3] = acc.w; }
  __syncthreads();
  if (threadIdx.x < 4 && wg * 4 + static_cast<int>(threadIdx.x) < words)
    bits[static_cast<size_t>(b) * words + wg * 4 + threadIdx.x] =
        (sh[0][threadIdx.x] | sh[1][threadIdx.x]) | (sh[2][threadIdx.x] | sh[3][threadIdx.x]);
}

// act_bits [n_img, words] -> count[f] (#images in which unit f fired), n_active[b] (#units fired in image b).
// utils.py:2047-2067.  grid.x over words (32 features each), 256 threads = 8 image lanes x 32 bit lanes.
static __global__ void activity_count_kernel(const uint32_t* __restrict__ bits, int n_img, int words, int F,
                                      float* __restrict__ count) {
  __shared__ int s[8][33];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int wd = blockIdx.x;
  int acc = 0;
  for (int b = g; b < n_img; b += 8) acc += (bits[static_cast<size_t>(b) * words + wd] >> lane) & 1u;
  s[g][lane] = acc;
  __syncthreads();
  if (g == 0) {
    int t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s[k][lane];
    const int f = wd * 32 + lane;
    if (f < F) count[f] = static_cast<float>(t);
  }
}
static __global__ void activity_per_image_kernel(const uint32_t* __restrict__ bits, int n_img, int words,
                                          int32_t* __restrict__ n_active, float* __restrict__ n_active_f) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= n_img) return;
  int acc = 0;
  for (int w = lane; w < words; w += 32) acc += __popc(bits[static_cast<size_t>(b) * words + w]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    if (n_active) n_active[b] = acc;
    if (n_active_f) n_active_f[b] = static_cast<float>(acc);
  }
}
// count[f] (global, after any all-reduce) -> dead mask, frequency, number of dead units.
static __global__ void activity_finalize_kernel(const float* __restrict__ count, int F, float n_images_global,
                                         uint8_t* __restrict__ dead, float* __restrict__ freq,
                                         float* __restrict__ n_dead_out) {
  __shared__ float s[32];
  float nd = 0.f;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const float c = count[f];
    const bool is_dead = c == 0.f;
    if (dead) dead[f] = is_dead ? 1 : 0;
    // 1 - mean(inactive) evaluated as the reference does: inactive-count / B, then 1 - that (utils.py:2056)
    if (freq) freq[f] = 1.f - (n_images_global - c) / n_images_global;
    nd += is_dead ? 1.f : 0.f;
  }
  const float r = block_sum(nd, s);
  if (threadIdx.x == 0 && n_dead_out) n_dead_out[0] = r;
}

// measure_inactive_units on a materialised tensor (API path): NCHW [B,F,HW] -> bits[b][f]; one warp per (b,f).
template <typename T>
static __global__ void activity_bits_nchw_kernel(const T* __restrict__ t, uint32_t* __restrict__ bits, int n_img, int F,
                                          int HW, int words) {
  const long long wid = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= static_cast<long long>(n_img) * F) return;
  const int b = static_cast<int>(wid / F), f = static_cast<int>(wid % F);
  const T* p = t + (static_cast<size_t>(b) * F + f) * HW;
  bool any = false;
  for (int i = lane; i < HW; i += 32) any |= (to_f32<T>(p[i]) != 0.f);
  any = __any_sync(0xffffffffu, any);
  if (lane == 0 && any) atomicOr(&bits[static_cast<size_t>(b) * words + (f >> 5)], 1u << (f & 31));
}
template <typename T>
static __global__ void activity_bits_rows_kernel(const T* __restrict__ t, uint32_t* __restrict__ bits, long long n_rows,
                                          int F, int words) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n_rows * words) return;
  const long long r = i / words;
  const int w = static_cast<int>(i % words);
  uint32_t m = 0;
  for (int j = 0; j < 32; ++j) {
    const int f = w * 32 + j;
    if (f < F && to_f32<T>(t[r * F + f]) != 0.f) m |= 1u << j;
  }
  bits[i] = m;
}

// One pass over a token-major bf16 tensor t [n_rows, F] (the 2-D call of measure_inactive_units, compute_ie.py:155: every
// token is a "sample"): count[f] += #rows with t[r, f] != 0 and n_active[r] = #units with t[r, f] != 0, without the bit
// matrix in between.  grid (ceil(F / 2048), row chunks of kActRows), 256 threads: a thread owns 8 consecutive units (one
// 16-byte load per row, the block reads 4 KB of contiguous row), eight counters in registers over the chunk's rows,
// one float atomicAdd per unit and chunk at the end (integers below 2^24: exact and order-independent); the row counts
// go through shared-memory counters.  count and n_active must be zero on entry.
constexpr int kActRows = 64;
static __global__ void __launch_bounds__(256)
activity_rows_fused_kernel(const uint4* __restrict__ t, long long n_rows, int F8, float* __restrict__ count,
                           int32_t* __restrict__ n_active) {
  __shared__ int s_row[kActRows];
  const int v = blockIdx.x * 256 + threadIdx.x;
  const long long r0 = static_cast<long long>(blockIdx.y) * kActRows;
  const int rows = static_cast<int>(min(static_cast<long long>(kActRows), n_rows - r0));
  if (threadIdx.x < kActRows) s_row[threadIdx.x] = 0;
  __syncthreads();
  int cnt[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) cnt[e] = 0;
  const bool in = v < F8;
#pragma unroll 4
  for (int r = 0; r < rows; ++r) {
    uint4 q = make_uint4(0, 0, 0, 0);
    if (in) q = __ldg(t + (r0 + r) * F8 + v);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    int pop = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int lo = (w[k] & 0x00007FFFu) != 0, hi = (w[k] & 0x7FFF0000u) != 0;   // +-0 is inactive, NaN / inf active
      cnt[2 * k] += lo;
      cnt[2 * k + 1] += hi;
      pop += lo + hi;
    }
    if (n_active) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pop += __shfl_xor_sync(0xffffffffu, pop, o);
      if ((threadIdx.x & 31) == 0 && pop) atomicAdd(&s_row[r], pop);
    }
  }
  if (in) {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (cnt[e]) atomicAdd(count + static_cast<size_t>(v) * 8 + e, static_cast<float>(cnt[e]));
  }
  if (n_active) {
    __syncthreads();
    if (threadIdx.x < rows && s_row[threadIdx.x]) atomicAdd(n_active + r0 + threadIdx.x, s_row[threadIdx.x]);
  }
}

// ------------------------------------------------------------------------------------------------ gradients
// Gradient assembly for the SaeMLP step (assemble_grads_kernel below).  All inputs are in "unscaled" units (see
// EpiDPre); s = 2/(T_global*C).
//   g_wdec[i] = s * sum_k P_wd[k][i]
//   g_wenc[f,c] = s * (sum_k P_we[k][f,c] - csum[f]*b_dec[c])     (G5 used x, not x - b_dec: rank-1 fix-up)
//   g_benc[f] = s * csum[f]
//   g_bdec[c] = s * (colsum(DIFF)[c] - sum_f csum[f] * W_enc[f,c])
// sum_splits_kernel: plain split-K reduction of svb_gemm_bf16.
static __global__ void sum_splits_kernel(const float* __restrict__ part, int splits, size_t n, float scale,
                                  float* __restrict__ out) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < splits; ++k) a += part[static_cast<size_t>(k) * n + i];
    out[i] = a * scale;
  }
}

// ------------------------------------------------------------------------------------------------ merged step kernels
// The small kernels around the GEMMs are launch-latency bound, so independent pieces share one launch: a block
// picks its job from its index ("roles").  Nothing here synchronises across blocks.

// Step prologue: bf16 shadow + folded bias of the encoder weight (role 0, one warp per row, see
// prep_encoder_kernel), bf16 shadow of the decoder weight (role 1), zeroing of the per-step accumulators (role 2)
// and, for the gated SAE, exp(r_mag) (role 3).
struct PrepArgs {
  const float* w_enc; const float* b_enc; const float* b_dec; bf16* w_enc_bf16; float* fold; float* dotw;
  const float* w_dec; bf16* w_dec_bf16;
  uint32_t* zero; unsigned long long n_zero;
  const float* r_mag; float* exp_r;
  int F, C;
  int nb_enc, nb_dec, nb_zero, nb_exp;
};
static __global__ void __launch_bounds__(256) prep_step_kernel(const PrepArgs a) {
  int blk = blockIdx.x;
  if (blk < a.nb_enc) {
    const int f = blk * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (f >= a.F) return;
    float acc = 0.f;
    for (int c = lane; c < a.C; c += 32) {
      const bf16 q = __float2bfloat16_rn(a.w_enc[static_cast<size_t>(f) * a.C + c]);
      a.w_enc_bf16[static_cast<size_t>(f) * a.C + c] = q;
      acc += __bfloat162float(q) * a.b_dec[c];
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (a.fold) a.fold[f] = (a.b_enc ? a.b_enc[f] : 0.f) - acc;
      if (a.dotw) a.dotw[f] = acc;
    }
    return;
  }
  blk -= a.nb_enc;
  if (blk < a.nb_dec) {
    const size_t n = static_cast<size_t>(a.F) * a.C;
    for (size_t i = (static_cast<size_t>(blk) * 256 + threadIdx.x) * 4; i < n; i += static_cast<size_t>(a.nb_dec) * 1024) {
      const float4 q = *reinterpret_cast<const float4*>(a.w_dec + i);  // F*C is a multiple of 64
      *reinterpret_cast<uint2*>(a.w_dec_bf16 + i) = make_uint2(pack_bf16x2(q.x, q.y), pack_bf16x2(q.z, q.w));
    }
    return;
  }
  blk -= a.nb_dec;
  if (blk < a.nb_zero) {
    for (size_t i = static_cast<size_t>(blk) * 256 + threadIdx.x; i < a.n_zero; i += static_cast<size_t>(a.nb_zero) * 256)
      a.zero[i] = 0u;
    return;
  }
  blk -= a.nb_zero;
  const int i = blk * 256 + threadIdx.x;
  if (i < a.F) a.exp_r[i] = expf(a.r_mag[i]);
}

// Gradient assembly after the two weight-gradient GEMMs (see the formulas above sum_splits_kernel):
//   role 0  g_wdec = s * sum_k P_wd[k]                       role 1  g_wenc = s * (sum_k P_we[k] - csum (x) b_dec)
//   role 2  g_benc = s * csum   (skipped when null)           role 3  vm[chunk] = csum[chunk] . W_enc_bf16[chunk, :]
//   role 4  activity count per unit                           role 5  active units per image
struct AssembleArgs {
  const float* P_wd; float* g_wdec; int s_wd;
  const float* P_we; float* g_wenc; int s_we;
  const float* csum; const float* b_dec; float* g_benc;
  const bf16* w_enc_bf16; float* vm; int vm_chunks;
  const uint32_t* act_bits; float* count; int32_t* n_active; float* nact_f; int n_img, words;
  int F, C;
  float s;
  int nb_wd, nb_w, nb_b, nb_vm, nb_cnt, nb_img;  // nb_wd = 0 / nb_w..nb_img = 0: that part is not run by this launch
};
static __global__ void __launch_bounds__(256) assemble_grads_kernel(const AssembleArgs a) {
  __shared__ int sh[8][33];
  int blk = blockIdx.x;
  const size_t n = static_cast<size_t>(a.F) * a.C;
  if (blk < a.nb_wd + a.nb_w) {
    const bool enc = blk >= a.nb_wd;
    if (enc) blk -= a.nb_wd;
    const float* part = enc ? a.P_we : a.P_wd;
    const int splits = enc ? a.s_we : a.s_wd;
    float* out = enc ? a.g_wenc : a.g_wdec;
    const int nb = enc ? a.nb_w : a.nb_wd;
    for (size_t i = (static_cast<size_t>(blk) * 256 + threadIdx.x) * 4; i < n; i += static_cast<size_t>(nb) * 1024) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < splits; ++k) {
        const float4 q = *reinterpret_cast<const float4*>(part + static_cast<size_t>(k) * n + i);
        acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
      }
      if (enc) {  // C % 8 == 0: the four elements share the row f
        const int f = static_cast<int>(i / a.C), c = static_cast<int>(i % a.C);
        const float cs = a.csum[f];
        const float4 b = *reinterpret_cast<const float4*>(a.b_dec + c);
        acc.x -= cs * b.x; acc.y -= cs * b.y; acc.z -= cs * b.z; acc.w -= cs * b.w;
      }
      *reinterpret_cast<float4*>(out + i) = make_float4(acc.x * a.s, acc.y * a.s, acc.z * a.s, acc.w * a.s);
    }
    return;
  }
  blk -= a.nb_wd + a.nb_w;
  if (blk < a.nb_b) {
    const int f = blk * 256 + threadIdx.x;
    if (a.g_benc && f < a.F) a.g_benc[f] = a.csum[f] * a.s;
    return;
  }
  blk -= a.nb_b;
  if (blk < a.nb_vm) {
    const int cblocks = (a.C + 255) / 256;
    const int chunk = blk / cblocks, c = (blk % cblocks) * 256 + threadIdx.x;
    const int per = (a.F + a.vm_chunks - 1) / a.vm_chunks;
    const int f0 = chunk * per, f1 = min(a.F, f0 + per);
    if (c >= a.C) return;
    float acc = 0.f;
    for (int f = f0; f < f1; ++f) acc += a.csum[f] * __bfloat162float(a.w_enc_bf16[static_cast<size_t>(f) * a.C + c]);
    a.vm[static_cast<size_t>(chunk) * a.C + c] = acc;
    return;
  }
  blk -= a.nb_vm;
  if (blk < a.nb_cnt) {  // utils.py:2047-2056: number of images in which unit f fired
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    int acc = 0;
    for (int b = g; b < a.n_img; b += 8) acc += (a.act_bits[static_cast<size_t>(b) * a.words + blk] >> lane) & 1u;
    sh[g][lane] = acc;
    __syncthreads();
    if (g == 0) {
      int t = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += sh[k][lane];
      const int f = blk * 32 + lane;
      if (f < a.F) a.count[f] = static_cast<float>(t);
    }
    return;
  }
  blk -= a.nb_cnt;
  {  // utils.py:2063-2067: active units per image
    const int b = blk * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= a.n_img) return;
    int acc = 0;
    for (int w = lane; w < a.words; w += 32) acc += __popc(a.act_bits[static_cast<size_t>(b) * a.words + w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      if (a.n_active) a.n_active[b] = acc;
      a.nact_f[b] = static_cast<float>(acc);
    }
  }
}

// Last kernel of the gradient half (ONE block): decoder-bias gradient, the loss partial sums and the per-channel
// statistics go into the flat reduction buffer.  sums section: [0] sum (d-x)^2, [1] sum |enc|, [2] sum aux^2,
// [3] sum Var(x), [4] sum Var(d), [5] sum active units, [6..7] zero.
struct TailArgs {
  const float* chan;  // [4][C]: sum diff, sum diff^2, min x, max x
  const float* vm; int vm_chunks; float* g_bdec; float s;
  const float* sq_part; int n_sq;
  const float* l1_part; int n_l1;
  const float* aux_part; int n_aux;  // null: no aux term
  const float* nact_f; int n_img;
  const float* var_part; int n_var_part; const float* rowvar; long long n_rows;
  float* flat; unsigned long long o_sums, o_chansq, o_max;
  int C;
};
static __global__ void __launch_bounds__(1024) grads_tail_kernel(const TailArgs a) {
  __shared__ float sh[32];
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < a.vm_chunks; ++k) acc += a.vm[static_cast<size_t>(k) * a.C + c];
    a.g_bdec[c] = (a.chan[c] - acc) * a.s;
    a.flat[a.o_chansq + c] = a.chan[a.C + c];
    a.flat[a.o_max + c] = a.chan[3 * a.C + c];
    a.flat[a.o_max + a.C + c] = -a.chan[2 * a.C + c];
  }
  auto total = [&](const float* p, long long n, long long stride, long long off) {
    float acc = 0.f;
    if (p)
      for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += p[i * stride + off];
    return block_sum(acc, sh);
  };
  const float sq = total(a.sq_part, a.n_sq, 1, 0);
  const float l1 = total(a.l1_part, a.n_l1, 1, 0);
  const float aux = total(a.aux_part, a.n_aux, 1, 0);
  const float na = total(a.nact_f, a.n_img, 1, 0);
  const float vx = a.n_rows > 0 ? total(a.rowvar, a.n_rows, 2, 0) : total(a.var_part, a.n_var_part, 2, 0);
  const float vd = a.n_rows > 0 ? total(a.rowvar, a.n_rows, 2, 1) : total(a.var_part, a.n_var_part, 2, 1);
  if (threadIdx.x == 0) {
    float* o = a.flat + a.o_sums;
    o[0] = sq; o[1] = l1; o[2] = aux; o[3] = vx; o[4] = vd; o[5] = na; o[6] = 0.f; o[7] = 0.f;
  }
}

// Step epilogue (ONE block, after any data-parallel all-reduce of the flat buffer): scalars of the step
// (utils.py:2470-2473, sparse_loss.py:4-21, utils.py:2012-2030,2063-2067) and the activity outputs
// (utils.py:2047-2060).
struct FinalizeArgs {
  const float* flat; unsigned long long o_sums, o_chansq, o_max, o_count;
  int C, F, expansion;
  float T_g, B_g, lambda;
  float* stats; uint8_t* dead; float* freq;
};
static __global__ void __launch_bounds__(1024) step_finalize_kernel(const FinalizeArgs a) {
  __shared__ float sh[32];
  float r = 0.f, nr = 0.f;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    const float rm = sqrtf(a.flat[a.o_chansq + c] / a.T_g);
    const float range = a.flat[a.o_max + c] + a.flat[a.o_max + a.C + c];  // max - min
    r += rm;
    nr += rm / range;
  }
  const float rs = block_sum(r, sh);
  const float nrs = block_sum(nr, sh);
  float nd = 0.f;
  for (int f = threadIdx.x; f < a.F; f += blockDim.x) {
    const float c = a.flat[a.o_count + f];
    const bool is_dead = c == 0.f;
    if (a.dead) a.dead[f] = is_dead ? 1 : 0;
    if (a.freq) a.freq[f] = 1.f - (a.B_g - c) / a.B_g;  // 1 - mean(inactive), evaluated like utils.py:2056
    nd += is_dead ? 1.f : 0.f;
  }
  const float nds = block_sum(nd, sh);
  if (threadIdx.x == 0 && a.stats) {
    const float* o = a.flat + a.o_sums;
    const float rec = o[0] / (a.T_g * a.C), l1 = o[1] / (a.T_g * a.F), aux = o[2] / (a.T_g * a.C);
    a.stats[SVB_STAT_REC] = rec;
    a.stats[SVB_STAT_L1] = l1;
    a.stats[SVB_STAT_AUX] = aux;
    a.stats[SVB_STAT_LOSS] = rec + a.lambda * l1 + aux;
    a.stats[SVB_STAT_RMSE] = rs / a.C;
    a.stats[SVB_STAT_NRMSE] = nrs / a.C;
    a.stats[SVB_STAT_VAR_EXPL] = 1.f - o[4] / o[3];
    a.stats[SVB_STAT_SPARSITY] = (o[5] / a.B_g) / (static_cast<float>(a.F) / a.expansion);
    a.stats[SVB_STAT_N_DEAD] = nds;
    for (int i = SVB_STAT_N_DEAD + 1; i < SVB_STATS_LEN; ++i) a.stats[i] = 0.f;   // callers hand in uninitialised blocks
  }
}

// ------------------------------------------------------------------------------------------------ optimiser
struct AdamCoef {
  float lr_over_bc1;    // lr / (1 - beta1^t)
  float inv_sqrt_bc2;   // 1 / sqrt(1 - beta2^t)
  float beta2, omb1, omb2, eps;   // beta2, 1 - beta1, 1 - beta2 (each rounded from the double), eps
  const float* dev;               // null, or device {lr_over_bc1, inv_sqrt_bc2} written by adam_step_coef_kernel
  __device__ __forceinline__ void resolve() {
    if (dev) { lr_over_bc1 = dev[0]; inv_sqrt_bc2 = dev[1]; }
  }
};
// Device-resident step counter (svb_opt_config::step_dev): t = ++(*step); out = {lr / (1 - b1^t), 1 / sqrt(1 - b2^t)}.
static __global__ void adam_step_coef_kernel(int32_t* step, double lr, double b1, double b2, float* out) {
  const int t = ++(*step);
  out[0] = static_cast<float>(lr / (1.0 - pow(b1, static_cast<double>(t))));
  out[1] = static_cast<float>(1.0 / sqrt(1.0 - pow(b2, static_cast<double>(t))));
}
__device__ __forceinline__ float adam_elem(float w, float g, float& m, float& v, const AdamCoef& k) {
  // torch.optim.Adam single-tensor update: m.lerp_(g, 1-b1); v = b2*v + (1-b2) g^2;
  // w -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
  m = m + (g - m) * k.omb1;
  v = v * k.beta2 + k.omb2 * g * g;
  const float denom = sqrtf(v) * k.inv_sqrt_bc2 + k.eps;
  return w - k.lr_over_bc1 * (m / denom);
}
static __global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, size_t n, AdamCoef k, bf16* __restrict__ w_bf16) {
  k.resolve();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float mi = m[i], vi = v[i];
    const float wn = adam_elem(w[i], g[i], mi, vi, k);
    w[i] = wn; m[i] = mi; v[i] = vi;
    if (w_bf16) w_bf16[i] = __float2bfloat16_rn(wn);
  }
}
// Plain Adam over several tensors in ONE launch: segment i owns blocks [blk0[i], blk0[i+1]).
struct AdamSeg { float* w; const float* g; float* m; float* v; unsigned long long n; };
struct AdamMultiArgs {
  AdamSeg seg[6];
  int blk0[7];
  int nseg;
};
static __global__ void __launch_bounds__(256) adam_multi_kernel(const AdamMultiArgs a, AdamCoef k) {
  k.resolve();
  int sgi = 0;
#pragma unroll
  for (int i = 1; i < 6; ++i)
    if (i < a.nseg && static_cast<int>(blockIdx.x) >= a.blk0[i]) sgi = i;
  const AdamSeg sg = a.seg[sgi];
  const int nb = a.blk0[sgi + 1] - a.blk0[sgi], lb = blockIdx.x - a.blk0[sgi];
  for (size_t i = static_cast<size_t>(lb) * 256 + threadIdx.x; i < sg.n; i += static_cast<size_t>(nb) * 256) {
    float mi = sg.m[i], vi = sg.v[i];
    const float wn = adam_elem(sg.w[i], sg.g[i], mi, vi, k);
    sg.w[i] = wn; sg.m[i] = mi; sg.v[i] = vi;
  }
}

// ConstrainedAdam on decoder.weight [C,F] (utils.py:65-81): per column f
//   w^ = w/|w|;  g' = g - (g.w^) w^;  Adam(g');  w /= |w|.      block = 16 columns x 64 row lanes (1024 threads):
// a warp covers two 64-byte row segments, and F/16 blocks keep most SMs busy for the three dependent passes.
constexpr int kCadamCols = 16, kCadamRows = 64;
static __global__ void __launch_bounds__(kCadamCols * kCadamRows)
constrained_adam_decoder_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, int C, int F, AdamCoef k) {
  k.resolve();
  __shared__ float s[kCadamRows][kCadamCols + 1];
  __shared__ float col_a[kCadamCols], col_b[kCadamCols];
  const int cl = threadIdx.x % kCadamCols, gl = threadIdx.x / kCadamCols;
  const int f = blockIdx.x * kCadamCols + cl;
  const bool ok = f < F;
  auto colsum = [&](float val, float* dst) {  // fixed order over the row lanes
    __syncthreads();
    s[gl][cl] = val;
    __syncthreads();
    if (gl == 0) {
      float t = 0.f;
      for (int q = 0; q < kCadamRows; ++q) t += s[q][cl];
      dst[cl] = t;
    }
    __syncthreads();
  };
  // pass 1: |w|^2 and g.w
  float nw = 0.f, gw = 0.f;
  if (ok)
    for (int c = gl; c < C; c += kCadamRows) {
      const float wi = w[static_cast<size_t>(c) * F + f], gi = g[static_cast<size_t>(c) * F + f];
      nw += wi * wi;
      gw += gi * wi;
    }
  colsum(nw, col_a);
  colsum(gw, col_b);
  const float norm = sqrtf(col_a[cl]);
  const float proj = col_b[cl] / norm;  // g . w^
  // pass 2: projected gradient (written back, as the reference mutates p.grad), Adam, new norm
  float nn = 0.f;
  if (ok)
    for (int c = gl; c < C; c += kCadamRows) {
      const size_t i = static_cast<size_t>(c) * F + f;
      const float wi = w[i];
      const float gp = g[i] - proj * (wi / norm);
      g[i] = gp;
      float mi = m[i], vi = v[i];
      const float wn = adam_elem(wi, gp, mi, vi, k);
      w[i] = wn; m[i] = mi; v[i] = vi;
      nn += wn * wn;
    }
  colsum(nn, col_a);
  const float inv = 1.f / sqrtf(col_a[cl]);
  if (ok)
    for (int c = gl; c < C; c += kCadamRows) {
      const size_t i = static_cast<size_t>(c) * F + f;
      w[i] *= inv;
    }
}
// Column renormalisation of a [C,F] matrix (sae_mlp.py:39,138).
static __global__ void renorm_columns_kernel(float* __restrict__ w, int C, int F) {
  __shared__ float s[8][33];
  __shared__ float col[32];
  const int lane = threadIdx.x & 31, gl = threadIdx.x >> 5;
  const int f = blockIdx.x * 32 + lane;
  float nw = 0.f;
  if (f < F)
    for (int c = gl; c < C; c += 8) { const float wi = w[static_cast<size_t>(c) * F + f]; nw += wi * wi; }
  s[gl][lane] = nw;
  __syncthreads();
  if (gl == 0) { float t = 0.f; for (int q = 0; q < 8; ++q) t += s[q][lane]; col[lane] = t; }
  __syncthreads();
  const float inv = 1.f / sqrtf(col[lane]);
  if (f < F)
    for (int c = gl; c < C; c += 8) w[static_cast<size_t>(c) * F + f] *= inv;
}

// Dead-unit scatter (sae_mlp.py:133-135,148-176): rows of W_enc / entries of b_enc / columns of W_dec of dead
// units are replaced and their Adam moments zeroed.
static __global__ void reinit_scatter_kernel(const uint8_t* __restrict__ dead, int F, int C, float* w_enc, float* b_enc,
                                      float* w_dec, const float* __restrict__ new_w_enc,
                                      const float* __restrict__ new_w_dec, float new_b_enc, float* m_we, float* v_we,
                                      float* m_be, float* v_be, float* m_wd, float* v_wd) {
  const size_t n = static_cast<size_t>(F) * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    {  // encoder layout [F,C]
      const int f = static_cast<int>(i / C);
      if (dead[f]) {
        w_enc[i] = new_w_enc[i];
        if (m_we) { m_we[i] = 0.f; v_we[i] = 0.f; }
        if (i % C == 0) {
          b_enc[f] = new_b_enc;
          if (m_be) { m_be[f] = 0.f; v_be[f] = 0.f; }
        }
      }
    }
    {  // decoder layout [C,F]
      const int f = static_cast<int>(i % F);
      if (dead[f]) {
        w_dec[i] = new_w_dec[i];
        if (m_wd) { m_wd[i] = 0.f; v_wd[i] = 0.f; }
      }
    }
  }
}

}  // namespace svb
