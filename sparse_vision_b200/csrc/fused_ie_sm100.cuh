// Fused node indirect effects of one SAE layer for C <= 256, C % 128 == 0 (sm_100a, SM pairs): the activations
// a = relu(x W_enc^T + fold) and the feature gradients G = g W_dec (nnsight_intervention_check.py:194-195) are produced
// tile by tile in TMEM and reduced on the spot, so neither [T, F] tensor is ever written (compute_ie.py:420-453,
// utils.py:2606-2637):
//   ie_feature[f] = mean_t | G[t,f] (avg[f, pos(t)] - a[t,f]) |
// and, for the SAE-error node (utils.py:2574-2602), the per-token sums  q[t] = sum_f a[t,f] G[t,f]  -- because
// sum_c g[t,c] dec[t,c] = q[t] + sum_c g[t,c] b_dec[c], the decoder GEMM of the un-fused path is not needed either.
//
// Transposed orientation, as in the fused backward (fused_bwd_sm100.cuh): FEATURES are the accumulator rows.  An SM pair
// owns 256 features (both weight tiles resident: W_enc K-major, W_dec MN-major, 64 KB each per CTA) and a strided set of
// 128-token blocks; per block two M = 256, N = 128 MMAs (each CTA streams half of the tokens of x and of g through a ring
// of 8 KB k-blocks) fill acc_a and acc_g, double-buffered in TMEM (2 x 256 columns).  An epilogue thread owns one feature:
// the token sum of |G (avg - a)| stays in a register for the whole kernel, and q is reduced across the warp's 32
// features with a 31-shuffle transpose-reduce per 32 tokens (one partial row per warp, summed by the error kernel).
// The running average avg[f, pos] is contiguous along positions, i.e. along the accumulator COLUMNS: read lane-per-feature
// it costs 32 cache lines per request (the first version: 7.8 kcycles per block for 2 kcycles of MMA).  Each warp
// therefore copies its [32 features x 32 positions] tile through shared memory with coalesced 4-byte cp.async (one
// line per request, issued one chunk ahead) and reads its own row back from a padded tile.
#pragma once
#include "gemm_host.cuh"
#include "epilogues.cuh"
#include "ptx_cluster.cuh"

namespace svb {

struct FusedIeParams {
  int T, C, F, HW;
  int slots;               // token-block slots per pair tile; grid = 2 * tiles_f * slots
  int tiles_f;             // ceil(F / 256) pair tiles
  const float* fold;       // [F] folded encoder bias
  const float* avg;        // [F, HW] fp32: running average of the encoder output (the reference's own layout)
  float* ie_part;          // [2 * slots][F]: sum_t |G (avg - a)| per slot and token half
  float* q_part;           // [tiles_f * 8][T]: sum over the warp's 32 features of a * G, row = (pair tile * 2 + rank) * 4 + lane quarter
#ifdef SVB_FIE_TRACE
  long long* trace;        // bring-up only: [grid][8] cycles spent waiting
#endif
};
#ifdef SVB_FIE_TRACE
#define FIE_WAIT(slot_, call) do { const long long t0_ = clock64(); call; tr[slot_] += clock64() - t0_; } while (0)
#define FIE_TRACE_DECL long long tr[8] = {0}
#else
#define FIE_WAIT(slot_, call) call
#define FIE_TRACE_DECL
#endif

namespace fie {
constexpr int kUnits = 8;                          // ring of 8 KB k-blocks: [64 tokens][64 channels]
constexpr uint32_t kUnit = 8192;
constexpr uint32_t kAvgWarp = 32 * 33 * 4;         // per epilogue warp: [32 features][32 positions + 1 pad] fp32
constexpr uint32_t kWeOff = 0, kWdOff = 65536, kRingOff = 131072, kAvgOff = kRingOff + kUnits * kUnit,
                   kBarOff = kAvgOff + 8 * kAvgWarp;
constexpr uint32_t kSmem = kBarOff + 256;
static_assert(kSmem <= kMaxDynSmem, "fused node-IE: shared memory budget");
struct Bars {
  uint64_t w_full, full[kUnits], empty[kUnits], acc_full[2], acc_empty[2];
  uint32_t tmem_ptr;
};
static_assert(sizeof(Bars) <= 256, "barrier block");
}  // namespace fie

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
fused_node_ie_kernel(const __grid_constant__ CUtensorMap tmWe, const __grid_constant__ CUtensorMap tmWd,
                     const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG, const FusedIeParams p) {
  using namespace fie;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Wes = smem + kWeOff;    // [C/64 k-blocks][128 f][64 c]            (A of the encoder MMA, K-major)
  uint8_t* Wds = smem + kWdOff;    // [C/64 k-blocks][2 atoms][64 c][64 f]    (A of the gradient MMA, MN-major)
  uint8_t* ring = smem + kRingOff;
  Bars* bar = reinterpret_cast<Bars*>(smem + kBarOff);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) / 32, 0);
  const int lane = static_cast<int>(threadIdx.x) % 32;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = static_cast<int>(blockIdx.x) >> 1;
  const int ptile = pair % p.tiles_f;
  const int slot = pair / p.tiles_f;
  const int f0 = ptile * 256 + static_cast<int>(rank) * 128;
  const int nblocks = (p.T + 127) / 128;
  const int n = (nblocks - slot + p.slots - 1) / p.slots;
  const int nkb = p.C / 64;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("svb: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmWe); tma_prefetch_desc(&tmWd); tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmG);
    mbar_init(&bar->w_full, 1);
    for (int s = 0; s < kUnits; ++s) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bar->acc_full[a], 1); mbar_init(&bar->acc_empty[a], 16); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem2_alloc(&bar->tmem_ptr, 512); tmem2_relinquish(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs, own halves)
    if (lane == 0) {
      const uint32_t w_full_l = mapa_u32(smem_u32(&bar->w_full), 0);
      if (leader) mbar_arrive_expect_tx(&bar->w_full, 2u * 2u * nkb * 16384u);
      for (int kb = 0; kb < nkb; ++kb) {
        tma2_load_2d(Wes + kb * 16384, &tmWe, w_full_l, kb * 64, f0);                                   // W_enc [F, C], box 64 x 128
        for (int j = 0; j < 2; ++j) tma2_load_2d(Wds + kb * 16384 + j * 8192, &tmWd, w_full_l, f0 + 64 * j, kb * 64);   // W_dec [C, F], box 64 x 64
      }
      uint32_t stage = 0, phase = 0;
      FIE_TRACE_DECL;
      for (int i = 0; i < n; ++i) {
        const int t0 = (slot + i * p.slots) * 128 + static_cast<int>(rank) * 64;   // this CTA's 64 tokens of the block
        for (int op = 0; op < 2; ++op) {                                            // x k-blocks, then g k-blocks
          for (int kb = 0; kb < nkb; ++kb) {
            FIE_WAIT(0, mbar_wait(&bar->empty[stage], phase ^ 1));
            if (leader) mbar_arrive_expect_tx(&bar->full[stage], 2u * kUnit);
            tma2_load_2d(ring + stage * kUnit, op == 0 ? &tmX : &tmG, mapa_u32(smem_u32(&bar->full[stage]), 0), kb * 64, t0);
            if (++stage == kUnits) { stage = 0; phase ^= 1; }
          }
        }
      }
#ifdef SVB_FIE_TRACE
      if (p.trace) p.trace[blockIdx.x * 8 + 0] = tr[0];
#endif
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread of the leader CTA)
    if (leader && lane == 0) {
      constexpr uint32_t idesc_a = make_idesc_bf16(256, 128, false, false);
      constexpr uint32_t idesc_g = make_idesc_bf16(256, 128, true, false);
      uint32_t stage = 0, phase = 0;
      FIE_TRACE_DECL;
      mbar_wait(&bar->w_full, 0);
      for (int i = 0; i < n; ++i) {
        const int a = i & 1;
        FIE_WAIT(1, mbar_wait(&bar->acc_empty[a], static_cast<uint32_t>((i >> 1) & 1) ^ 1u));
        tc_fence_after();
        for (int op = 0; op < 2; ++op) {
          const uint32_t d_tmem = tmem_base + a * 256 + op * 128;
          for (int kb = 0; kb < nkb; ++kb) {
            FIE_WAIT(2, mbar_wait(&bar->full[stage], phase));
            tc_fence_after();
            const uint32_t b_base = smem_u32(ring + stage * kUnit);
            const uint32_t a_base = smem_u32((op == 0 ? Wes : Wds) + kb * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adesc = op == 0 ? make_smem_desc_sw128(a_base + k * 32, 16, 1024)
                                             : make_smem_desc_sw128(a_base + k * 2048, 8192, 1024);
              umma2_f16(d_tmem, adesc, make_smem_desc_sw128(b_base + k * 32, 16, 1024), op == 0 ? idesc_a : idesc_g,
                        (kb | k) != 0 ? 1u : 0u);
            }
            umma2_commit_both(&bar->empty[stage]);
            if (++stage == kUnits) { stage = 0; phase ^= 1; }
          }
        }
        umma2_commit_both(&bar->acc_full[a]);
      }
#ifdef SVB_FIE_TRACE
      if (p.trace) { p.trace[blockIdx.x * 8 + 1] = tr[1]; p.trace[blockIdx.x * 8 + 2] = tr[2]; }
#endif
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 4 lane quarters x 2 token halves, both CTAs
    const int ew = warp - 2, wq = warp % 4, h = ew / 4;
    const int f = f0 + wq * 32 + lane;
    const bool f_ok = f < p.F;
    const float nfold = f_ok ? __ldg(p.fold + f) : 0.f;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t acc_empty_l[2] = {mapa_u32(smem_u32(&bar->acc_empty[0]), 0), mapa_u32(smem_u32(&bar->acc_empty[1]), 0)};
    float* q_row = p.q_part + (static_cast<size_t>(ptile * 2 + static_cast<int>(rank)) * 4 + wq) * p.T;
    float* avg_s = reinterpret_cast<float*>(smem + kAvgOff + ew * kAvgWarp);
    const int fw0 = f0 + wq * 32;                          // first feature of this warp
    // cp.async of the warp's avg tile for the 32 tokens from t0 on: request r copies feature fw0 + r, lane = position
    auto fetch_avg = [&](long long t0) {
      int pos = static_cast<int>((t0 + lane) % p.HW);
      const uint32_t dst = smem_u32(avg_s + lane);
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const int fr = fw0 + r < p.F ? fw0 + r : p.F - 1;  // rows of features >= F are never used (G = 0 there)
        const float* src = p.avg + static_cast<size_t>(fr) * p.HW + pos;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + r * 33 * 4), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float ie = 0.f;
#ifdef SVB_FIE_TRACE
    long long tr[8] = {0};
    const long long tr_start = clock64();
#endif
    if (n > 0) fetch_avg(static_cast<long long>(slot) * 128 + h * 64);
    for (int i = 0; i < n; ++i) {
      const int a = i & 1;
      const long long tb = static_cast<long long>(slot + i * p.slots) * 128 + h * 64;
      FIE_WAIT(3, mbar_wait(&bar->acc_full[a], static_cast<uint32_t>((i >> 1) & 1)));
      tc_fence_after();
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const long long t0 = tb + ci * 32;
        float va[32], vg[32];
        tmem_ld_32x32(lane_base + a * 256 + h * 64 + ci * 32, va);
        tmem_ld_32x32(lane_base + a * 256 + 128 + h * 64 + ci * 32, vg);
        // this lane's row of the staged average tile, then the next chunk's tile is requested
        float av[32];
        FIE_WAIT(4, asm volatile("cp.async.wait_group 0;" ::: "memory"); __syncwarp());
#pragma unroll
        for (int j = 0; j < 32; ++j) av[j] = avg_s[lane * 33 + j];
        __syncwarp();
        if (ci == 0) fetch_avg(t0 + 32);
        else if (i + 1 < n) fetch_avg(static_cast<long long>(slot + (i + 1) * p.slots) * 128 + h * 64);
        FIE_WAIT(5, tmem_ld_wait());
        if (ci == 1) {                                   // both chunks of this block are in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc_empty_l[a]);
        }
        float q[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float act = f_ok ? fmaxf(va[j] + nfold, 0.f) : 0.f;
          const float gg = vg[j];                         // 0 for tokens >= T (g rows are zero-filled) and features >= F
          ie += fabsf(gg * (av[j] - act));
          q[j] = act * gg;
        }
        // transpose-reduce: afterwards lane l holds sum over the warp's 32 features of q[l]
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
          const bool up = (lane & o) != 0;
#pragma unroll
          for (int j = 0; j < o; ++j) {
            const float keep = up ? q[j + o] : q[j];
            const float send = up ? q[j] : q[j + o];
            q[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
        if (t0 + lane < p.T) q_row[t0 + lane] = q[0];
      }
    }
#ifdef SVB_FIE_TRACE
    if (p.trace && ew == 0 && lane == 0) {
      for (int q_ = 3; q_ < 6; ++q_) p.trace[blockIdx.x * 8 + q_] = tr[q_];
      p.trace[blockIdx.x * 8 + 6] = clock64() - tr_start;
    }
#endif
    if (f_ok) p.ie_part[static_cast<size_t>(2 * slot + h) * p.F + f] = ie;
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem2_dealloc(tmem_base, 512);
}

#ifdef SVB_FIE_TRACE
inline long long*& fused_ie_trace_ptr() {   // bring-up: device buffer [grid][8] of wait cycles
  static long long* p = nullptr;
  return p;
}
#endif
inline int fused_ie_slots(long long T, int F, int max_ctas = 0) {
  const int pairs = (max_ctas > 0 ? max_ctas : device_sm_count()) / 2;
  const int tiles_f = (F + 255) / 256;
  if (tiles_f > pairs) return 0;
  long long s = pairs / tiles_f;
  const long long nblocks = (T + 127) / 128;
  return static_cast<int>(s < nblocks ? s : nblocks);
}
inline bool fused_ie_supported(long long T, int C, int F, int max_ctas = 0) {
  return C % 128 == 0 && C >= 128 && C <= 256 && F % 8 == 0 && T > 0 && T < (1ll << 31) - 256 && fused_ie_slots(T, F, max_ctas) >= 1;
}
inline int fused_ie_qrows(int F) { return ((F + 255) / 256) * 8; }

// x, g: bf16 token matrices [T, C] (row-major); w_enc bf16 [F, C], w_dec bf16 [C, F]; avg fp32 [F, HW].
// ie_part: [2 * slots][F], q_part: [fused_ie_qrows(F)][T].
inline int launch_fused_node_ie(cudaStream_t stream, const void* x, const void* g, const void* w_enc, const void* w_dec,
                                const float* fold, const float* avg, int T, int C, int F, int HW, float* ie_part, float* q_part,
                                int max_ctas = 0) {
  if (!fused_ie_supported(T, C, F, max_ctas)) return -2;
  CUtensorMap tmWe, tmWd, tmX, tmG;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmWe, w_enc, F, C, C, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmWd, w_dec, C, F, F, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmX, x, T, C, C, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmG, g, T, C, C, 64))) return rc;
  FusedIeParams p;
  p.T = T; p.C = C; p.F = F; p.HW = HW;
  p.tiles_f = (F + 255) / 256;
  p.slots = fused_ie_slots(T, F, max_ctas);
  p.fold = fold; p.avg = avg; p.ie_part = ie_part; p.q_part = q_part;
#ifdef SVB_FIE_TRACE
  p.trace = fused_ie_trace_ptr();
#endif
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return -4;
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(fused_node_ie_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fie::kSmem) != cudaSuccess) return -4;
    configured[dev] = true;
  }
  (fused_node_ie_kernel<<<2 * p.tiles_f * p.slots, 320, fie::kSmem, stream>>>(tmWe, tmWd, tmX, tmG, p), svb::count_launch());
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace svb
