// Stand-alone bring-up probe for the fused encoder -> decoder kernel planned in DESIGN.md section 8 item 1 (NOT part of
// the shipped library): one CTA per 128-token tile keeps X resident, walks the feature tiles, and for each of them
//   GEMM1  acc1[128 x 256] = X[128 x 256] * W_enc[f-tile]^T        (TMEM columns 0..255)
//   epilogue: relu, bf16, written ONCE into shared memory in the UMMA K-major / 128B-swizzle layout; from there the
//             tile is TMA-stored to E AND consumed as the A operand of
//   GEMM2  acc2[128 x 256] += E_tile[128 x 256] * W_dec[:, f-tile]^T (TMEM columns 256..511)
// so E is written but never re-read from HBM; D leaves once per token tile.  C = 256 only; no biases, masks, statistics
// (the product epilogues are not reproduced here).  Weight tiles are streamed from L2 through a 3-stage ring without
// multicast, so this measures the un-shared weight traffic the design note warns about.
//   fused_fwd_probe check        exact integer-valued check (T = 384, F = 768) against naive kernels
//   fused_fwd_probe perf         cfg2 size (T = 200704, F = 2048), time per call
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../sparse_vision_b200/csrc/gemm_host.cuh"

using namespace svb;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

namespace {

constexpr int kC = 256;                 // channels = K of GEMM1 = N of GEMM2
constexpr int kFT = 256;                // feature tile = N of GEMM1 = K chunk of GEMM2
#ifndef PROBE_STAGES
#define PROBE_STAGES 3
#endif
constexpr int kStages = PROBE_STAGES;
constexpr uint32_t kTileA = 128 * 64 * 2;      // 16 KB: one 128-row x 64-column K-major block
constexpr uint32_t kTileB = 256 * 64 * 2;      // 32 KB: one 256-row x 64-column K-major block
constexpr uint32_t kXsOff = 0, kEsOff = 4 * kTileA, kRingOff = 8 * kTileA, kBarOff = kRingOff + kStages * kTileB;
constexpr uint32_t kSmem = kBarOff + 256;
static_assert(kSmem <= kMaxDynSmem, "shared memory budget");

struct Bars {
  uint64_t x_full, x_empty, full[kStages], empty[kStages], acc1_full, acc1_empty, es_full, es_empty, acc2_full, acc2_empty;
  uint32_t tmem_ptr;
};

__global__ void __launch_bounds__(320, 1)
fused_fwd_probe_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWe,
                       const __grid_constant__ CUtensorMap tmWd, const __grid_constant__ CUtensorMap tmE,
                       const __grid_constant__ CUtensorMap tmD, int T, int F) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Xs = smem + kXsOff;
  uint8_t* Es = smem + kEsOff;
  uint8_t* ring = smem + kRingOff;
  Bars* bar = reinterpret_cast<Bars*>(smem + kBarOff);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) / 32, 0);
  const int lane = static_cast<int>(threadIdx.x) % 32;
  const int tiles_m = (T + 127) / 128, NF = F / kFT;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) { printf("probe: smem not 1024-aligned\n"); __trap(); }
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmWe); tma_prefetch_desc(&tmWd);
    tma_prefetch_desc(&tmE); tma_prefetch_desc(&tmD);
    mbar_init(&bar->x_full, 1); mbar_init(&bar->x_empty, 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
    mbar_init(&bar->acc1_full, 1); mbar_init(&bar->acc1_empty, 8);
    mbar_init(&bar->es_full, 8); mbar_init(&bar->es_empty, 1);
    mbar_init(&bar->acc2_full, 1); mbar_init(&bar->acc2_empty, 8);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&bar->tmem_ptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, px = 0;
      auto load_w = [&](const CUtensorMap* tm, int col, int row) {
        mbar_wait(&bar->empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&bar->full[stage], kTileB);
        tma_load_2d(ring + stage * kTileB, tm, &bar->full[stage], col, row);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      };
      for (int t = blockIdx.x; t < tiles_m; t += gridDim.x) {
        mbar_wait(&bar->x_empty, px ^ 1);
        mbar_arrive_expect_tx(&bar->x_full, 4 * kTileA);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(Xs + kb * kTileA, &tmX, &bar->x_full, kb * 64, t * 128);
        px ^= 1;
        for (int j = 0; j < NF; ++j) {
          for (int kb = 0; kb < 4; ++kb) load_w(&tmWe, kb * 64, j * kFT);                    // W_enc [F, C]
          if (j >= 1) for (int kb = 0; kb < 4; ++kb) load_w(&tmWd, (j - 1) * kFT + kb * 64, 0);   // W_dec [C, F]
        }
        for (int kb = 0; kb < 4; ++kb) load_w(&tmWd, (NF - 1) * kFT + kb * 64, 0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 256, false, false);
      uint32_t stage = 0, phase = 0, px = 0, p_a1e = 0, p_esf = 0, p_a2e = 0;
      const uint32_t acc1 = tmem_base, acc2 = tmem_base + 256;
      auto mma_blocks = [&](uint32_t d_tmem, const uint8_t* a_tile, bool first_accumulates) {
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(&bar->full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(a_tile + kb * kTileA), b_base = smem_u32(ring + stage * kTileB);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(d_tmem, make_smem_desc_sw128(a_base + k * 32, 16, 1024), make_smem_desc_sw128(b_base + k * 32, 16, 1024),
                     idesc, (first_accumulates || (kb | k) != 0) ? 1u : 0u);
          umma_commit(&bar->empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      };
      auto gemm2 = [&](int jj) {
        mbar_wait(&bar->es_full, p_esf); p_esf ^= 1;
        if (jj == 0) { mbar_wait(&bar->acc2_empty, p_a2e ^ 1); p_a2e ^= 1; }
        tc_fence_after();
        mma_blocks(acc2, Es, jj != 0);
        umma_commit(&bar->es_empty);
      };
      for (int t = blockIdx.x; t < tiles_m; t += gridDim.x) {
        mbar_wait(&bar->x_full, px); px ^= 1;
        for (int j = 0; j < NF; ++j) {
          mbar_wait(&bar->acc1_empty, p_a1e ^ 1); p_a1e ^= 1;
          tc_fence_after();
          mma_blocks(acc1, Xs, false);
          umma_commit(&bar->acc1_full);
          if (j == NF - 1) umma_commit(&bar->x_empty);
          if (j >= 1) gemm2(j - 1);
        }
        gemm2(NF - 1);
        umma_commit(&bar->acc2_full);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 8 warps = 4 lane quarters x 2 column halves
    const int ew = warp - 2, wq = warp % 4, cgroup = ew / 4;
    const int r = wq * 32 + lane;                       // row inside the tile
    uint32_t p_a1f = 0, p_ese = 0, p_a2f = 0;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    for (int t = blockIdx.x; t < tiles_m; t += gridDim.x) {
      const int m0 = t * 128;
      for (int j = 0; j <= NF; ++j) {                   // j == NF: the D tile of this token tile
        const bool is_d = j == NF;
        if (!is_d) { mbar_wait(&bar->acc1_full, p_a1f); p_a1f ^= 1; }
        else { mbar_wait(&bar->acc2_full, p_a2f); p_a2f ^= 1; }
        tc_fence_after();
        uint32_t pk[4][16];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          float v[32];
          tmem_ld_32x32(lane_base + (is_d ? 256u : 0u) + (cgroup * 4 + ci) * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            float a = v[2 * q], b = v[2 * q + 1];
            if (!is_d) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
            pk[ci][q] = pack_bf16x2(a, b);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(is_d ? &bar->acc2_empty : &bar->acc1_empty);
        // Es is free once GEMM2 of the previous feature tile has read it (not needed for j == 0: the previous user
        // was this warp's own D store) and once this warp's own TMA stores have finished reading it
        if (j >= 1) { mbar_wait(&bar->es_empty, p_ese); p_ese ^= 1; }
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = cgroup * 4 + ci;                // 32-column chunk of the tile
          uint8_t* row = Es + (c >> 1) * kTileA + r * 128;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int p16 = (c & 1) * 4 + i;
            *reinterpret_cast<uint4*>(row + ((p16 ^ (r & 7)) << 4)) =
                make_uint4(pk[ci][4 * i], pk[ci][4 * i + 1], pk[ci][4 * i + 2], pk[ci][4 * i + 3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (!is_d) mbar_arrive(&bar->es_full);        // GEMM2 may read the tile (all 8 warps arrive)
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const int kbi = cgroup * 2 + s;
            const uint8_t* src = Es + kbi * kTileA + wq * 4096;
            if (!is_d) tma_store_2d(&tmE, src, j * kFT + kbi * 64, m0 + wq * 32);
            else tma_store_2d(&tmD, src, kbi * 64, m0 + wq * 32);
          }
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait<0>();
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// naive references (exact for the integer-valued test inputs)
__global__ void ref_enc_kernel(const __nv_bfloat16* X, const __nv_bfloat16* We, __nv_bfloat16* E, int T, int F) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (f >= F || t >= T) return;
  float a = 0.f;
  for (int c = 0; c < kC; ++c) a += __bfloat162float(X[(size_t)t * kC + c]) * __bfloat162float(We[(size_t)f * kC + c]);
  E[(size_t)t * F + f] = __float2bfloat16_rn(fmaxf(a, 0.f));
}
__global__ void ref_dec_kernel(const __nv_bfloat16* E, const __nv_bfloat16* Wd, __nv_bfloat16* D, int T, int F) {
  const int c = threadIdx.x, t = blockIdx.x;
  float a = 0.f;
  for (int f = 0; f < F; ++f) a += __bfloat162float(E[(size_t)t * F + f]) * __bfloat162float(Wd[(size_t)c * F + f]);
  D[(size_t)t * kC + c] = __float2bfloat16_rn(a);
}

uint32_t rng = 12345u;
int rnd3() { rng = rng * 1664525u + 1013904223u; return (int)((rng >> 24) % 3) - 1; }   // -1, 0, 1

int run(int T, int F, bool check, int iters) {
  std::vector<__nv_bfloat16> hX((size_t)T * kC), hWe((size_t)F * kC), hWd((size_t)kC * F);
  for (auto& v : hX) v = __float2bfloat16((float)rnd3());
  for (auto& v : hWe) v = __float2bfloat16((float)rnd3());
  for (auto& v : hWd) v = __float2bfloat16((float)rnd3());   // all sums are integers < 2^24: fp32 exact, then one RN to bf16
  __nv_bfloat16 *dX, *dWe, *dWd, *dE, *dD, *rE, *rD;
  CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dWe, hWe.size() * 2)); CK(cudaMalloc(&dWd, hWd.size() * 2));
  CK(cudaMalloc(&dE, (size_t)T * F * 2)); CK(cudaMalloc(&dD, (size_t)T * kC * 2));
  CK(cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dWe, hWe.data(), hWe.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dWd, hWd.data(), hWd.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dE, 0xFF, (size_t)T * F * 2)); CK(cudaMemset(dD, 0xFF, (size_t)T * kC * 2));
  CUtensorMap tmX, tmWe, tmWd, tmE, tmD;
  if (make_tmap_bf16_2d(&tmX, dX, T, kC, kC, 128) || make_tmap_bf16_2d(&tmWe, dWe, F, kC, kC, 256) ||
      make_tmap_bf16_2d(&tmWd, dWd, kC, F, F, 256) || make_store_tmap_bf16(&tmE, dE, T, F, F) ||
      make_store_tmap_bf16(&tmD, dD, T, kC, kC)) { printf("tensor map failed\n"); return 2; }
  CK(cudaFuncSetAttribute(fused_fwd_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  const int tiles_m = (T + 127) / 128, sms = device_sm_count();
  const int grid = tiles_m < sms ? tiles_m : sms;
  fused_fwd_probe_kernel<<<grid, 320, kSmem>>>(tmX, tmWe, tmWd, tmE, tmD, T, F);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  if (check) {
    CK(cudaMalloc(&rE, (size_t)T * F * 2)); CK(cudaMalloc(&rD, (size_t)T * kC * 2));
    ref_enc_kernel<<<dim3((F + 255) / 256, T), 256>>>(dX, dWe, rE, T, F);
    ref_dec_kernel<<<T, kC>>>(rE, dWd, rD, T, F);
    CK(cudaDeviceSynchronize());
    std::vector<uint16_t> a((size_t)T * F), b((size_t)T * F), c((size_t)T * kC), d((size_t)T * kC);
    CK(cudaMemcpy(a.data(), dE, a.size() * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), rE, b.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c.data(), dD, c.size() * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(d.data(), rD, d.size() * 2, cudaMemcpyDeviceToHost));
    size_t be = 0, bd = 0;
    for (size_t i = 0; i < a.size(); ++i) if (a[i] != b[i] && ++be <= 5) printf("E mismatch at (%zu,%zu): %04x vs %04x\n", i / F, i % F, a[i], b[i]);
    for (size_t i = 0; i < c.size(); ++i) if (c[i] != d[i] && ++bd <= 5) printf("D mismatch at (%zu,%zu): %04x vs %04x\n", i / kC, i % kC, c[i], d[i]);
    printf("check T=%d F=%d: E mismatches %zu / %zu, D mismatches %zu / %zu -> %s\n", T, F, be, a.size(), bd, c.size(),
           (be | bd) ? "FAIL" : "PASS");
    return (be | bd) ? 1 : 0;
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) fused_fwd_probe_kernel<<<grid, 320, kSmem>>>(tmX, tmWe, tmWd, tmE, tmD, T, F);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  printf("perf T=%d F=%d: %.3f ms per call = %.0f TFLOP/s over both GEMMs (un-fused enc + dec in the step: ~0.42 ms)\n", T, F,
         ms, 4.0 * T * kC * F / (ms * 1e-3) * 1e-12);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  const char* mode = argc > 1 ? argv[1] : "check";
  if (!strcmp(mode, "check")) {
    int rc = run(384, 768, true, 0);
    rc |= run(128 * 300, 512, true, 0);      // more tiles than SMs: the persistent loop and every barrier phase wrap
    return rc;
  }
  return run(200704, 2048, false, 10);
}
