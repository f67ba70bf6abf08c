// Indirect-effect reductions (utils.py:2574-2660).  HBM-bound: every input element is read exactly once with 16-byte
// coalesced loads, the [F,H,W] average is read through a token-major transposed copy that stays L2-resident across
// images, and the per-feature sums are reduced with a fixed two-stage order (no floating-point atomics), so top-k
// feature sets are stable run to run and across GPU counts.
#pragma once
#include "kernels_misc.cuh"

namespace svb {

// avg [F, HW] -> avgT [HW, F]  (what reshape_encoder_output_average + rearrange produce per image, utils.py:2776-2782,
// without materialising the batch_size-fold repeat).
static __global__ void transpose_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? in[static_cast<size_t>(r) * Cc + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < Cc && r < R) out[static_cast<size_t>(c) * R + r] = tile[threadIdx.x][i];
  }
}

template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int kN = 4;
  __device__ static void load(const float* p, float (&o)[8]) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
  }
};
template <> struct Vec16<bf16> {
  static constexpr int kN = 8;
  __device__ static void load(const bf16* p, float (&o)[8]) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    o[0] = bf16lo(q.x); o[1] = bf16hi(q.x); o[2] = bf16lo(q.y); o[3] = bf16hi(q.y);
    o[4] = bf16lo(q.z); o[5] = bf16hi(q.z); o[6] = bf16lo(q.w); o[7] = bf16hi(q.w);
  }
};

// compute_ie_channel_wise (utils.py:2606-2637):  partial[chunk][f] = sum_{t in chunk} | g[t,f] * (avg[f, t%HW] - a[t,f]) |
//   a, g : [T, F] token-major;  avgT : [HW, F] fp32.
//   grid (ceil(F / (32*kN)), chunks), 256 threads: lane owns kN consecutive features, warp w owns rows w, w+8, ...
// Requires F % kN == 0 (16-byte loads).
template <typename T, int UNROLL>
static __global__ void __launch_bounds__(256)
ie_channelwise_kernel(const T* __restrict__ a, const T* __restrict__ g, const float* __restrict__ avgT, long long Tn,
                      int HW, int F, float* __restrict__ partial) {
  constexpr int V = Vec16<T>::kN;
  __shared__ float s[8][32 * V + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int f0 = (blockIdx.x * 32 + lane) * V;
  const long long rows_per = (Tn + gridDim.y - 1) / gridDim.y;
  const long long r_begin = blockIdx.y * rows_per;
  const long long r_end = r_begin + rows_per < Tn ? r_begin + rows_per : Tn;
  float acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = 0.f;
  if (f0 < F) {
    long long r = r_begin + w;
    for (; r + 8 * (UNROLL - 1) < r_end; r += 8 * UNROLL) {
      float av[UNROLL][8], gv[UNROLL][8], mv[UNROLL][8];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const long long rr = r + 8 * u;
        Vec16<T>::load(a + rr * F + f0, av[u]);
        Vec16<T>::load(g + rr * F + f0, gv[u]);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int hw = static_cast<int>((r + 8 * u) % HW);
        const float* m = avgT + static_cast<size_t>(hw) * F + f0;
#pragma unroll
        for (int q = 0; q < V; q += 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(m + q));
          mv[u][q] = t.x; mv[u][q + 1] = t.y; mv[u][q + 2] = t.z; mv[u][q + 3] = t.w;
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += fabsf(gv[u][k] * (mv[u][k] - av[u][k]));
    }
    for (; r < r_end; r += 8) {
      float av[8], gv[8];
      Vec16<T>::load(a + r * F + f0, av);
      Vec16<T>::load(g + r * F + f0, gv);
      const int hw = static_cast<int>(r % HW);
      const float* m = avgT + static_cast<size_t>(hw) * F + f0;
#pragma unroll
      for (int k = 0; k < V; ++k) acc[k] += fabsf(gv[k] * (__ldg(m + k) - av[k]));
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) s[w][lane * V + k] = acc[k];
  __syncthreads();
  for (int j = threadIdx.x; j < 32 * V; j += 256) {
    const int f = blockIdx.x * 32 * V + j;
    if (f < F) {
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += s[q][j];
      partial[static_cast<size_t>(blockIdx.y) * F + f] = t;
    }
  }
}

// compute_ie_all_channels (utils.py:2574-2602) on NCHW tensors:
//   partial[block] = sum_{pixels of block} | sum_c g[b,c,p] * (avg[c,p] - err[b,c,p]) |
//   block (32 pixels, 8 channel lanes).
template <typename T>
static __global__ void ie_allchannels_nchw_kernel(const T* __restrict__ err, const T* __restrict__ g,
                                           const float* __restrict__ avg, long long n_pix, int C, int HW,
                                           float* __restrict__ partial) {
  __shared__ float s[8][33];
  const long long px = blockIdx.x * 32LL + threadIdx.x;
  float acc = 0.f;
  if (px < n_pix) {
    const long long b = px / HW;
    const int p = static_cast<int>(px % HW);
    const size_t base = static_cast<size_t>(b) * C * HW + p;
    for (int c = threadIdx.y; c < C; c += 8) {
      const size_t i = base + static_cast<size_t>(c) * HW;
      acc += to_f32<T>(g[i]) * (avg[static_cast<size_t>(c) * HW + p] - to_f32<T>(err[i]));
    }
  }
  s[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += s[q][threadIdx.x];
    t = px < n_pix ? fabsf(t) : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
  }
}

// Token-major variant used by the fused node-IE layer: q, g : [T, C] bf16;  avgT : [HW, C] fp32;
//   value_t = | sum_c g[t,c] * (avgT[t%HW, c] - sign * q[t,c]) |        one warp per token, 8 tokens per block.
static __global__ void ie_allchannels_tokens_kernel(const bf16* __restrict__ q, const bf16* __restrict__ g,
                                             const float* __restrict__ avgT, long long Tn, int C, int HW, float sign,
                                             float* __restrict__ partial) {
  __shared__ float s[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long t = blockIdx.x * 8LL + w;
  float acc = 0.f;
  if (t < Tn) {
    const int hw = static_cast<int>(t % HW);
    for (int c = lane * 8; c < C; c += 256) {
      float qv[8], gv[8];
      Vec16<bf16>::load(q + t * C + c, qv);
      Vec16<bf16>::load(g + t * C + c, gv);
      const float* m = avgT + static_cast<size_t>(hw) * C + c;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += gv[k] * (__ldg(m + k) - sign * qv[k]);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) s[w] = t < Tn ? fabsf(acc) : 0.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += s[k];
    partial[blockIdx.x] = r;
  }
}

// SAE-error node of the FUSED node-IE layer (fused_ie_sm100.cuh): the decoder output is never formed, because
//   sum_c g[t,c] (err_avg[pos,c] - err[t,c]),  err = x - (a W_dec^T + b_dec)
//     = sum_c g[t,c] (err_avg[pos,c] - x[t,c] + b_dec[c])  +  sum_f a[t,f] G[t,f]
// and the fused kernel leaves the second sum as `nrows` partial rows qrows[r][t].  One warp per token, 8 tokens per block.
static __global__ void ie_error_tokens_fused_kernel(const bf16* __restrict__ x, const bf16* __restrict__ g,
                                                    const float* __restrict__ avgT, const float* __restrict__ bias,
                                                    const float* __restrict__ qrows, int nrows, long long Tn, int C, int HW,
                                                    float* __restrict__ partial) {
  __shared__ float s[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long t = blockIdx.x * 8LL + w;
  float acc = 0.f;
  if (t < Tn) {
    const int hw = static_cast<int>(t % HW);
    for (int c = lane * 8; c < C; c += 256) {
      float xv[8], gv[8];
      Vec16<bf16>::load(x + t * C + c, xv);
      Vec16<bf16>::load(g + t * C + c, gv);
      const float* m = avgT + static_cast<size_t>(hw) * C + c;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += gv[k] * (__ldg(m + k) - xv[k] + __ldg(bias + c + k));
    }
    for (int r = lane; r < nrows; r += 32) acc += __ldg(qrows + static_cast<size_t>(r) * Tn + t);
  }
  acc = warp_sum(acc);
  if (lane == 0) s[w] = t < Tn ? fabsf(acc) : 0.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += s[k];
    partial[blockIdx.x] = r;
  }
}

}  // namespace svb
