// Two-CTA (cta_group::2) variant of the B-stationary GEMM for K <= 256:  D[M,N] = A[M,K] * B[N,K]^T.
// (Round 1 measured this kernel SLOWER than the single-CTA one and parked it under tools/; the cause was not the
// pairing but the `.release.cluster` qualifier on the epilogue's remote mbarrier arrives -- a MEMBAR.ALL.GPU per
// arrive, see ptx_cluster.cuh.  With default-semantics arrives the output-bound K = 256 probe runs at 0.154 ms
// against 0.173-0.194 ms single-CTA, and the SaeMLP encoder GEMM uses it.)
//
// A cluster of two CTAs (an SM pair) owns one 256-column N tile and walks 256-row "pair tiles" of M; CTA r of the
// pair computes rows [256*tp + 128*r, +128).  One thread of the leader CTA issues tcgen05.mma.cta_group::2 with
// M = 256: each CTA contributes its own 128 x K slice of A and HALF of the weight tile (128 of the 256 N rows), and the
// hardware shares the halves between the two SMs.  Per CTA that is
//   * 64 KB instead of 128 KB of resident weights (room for a deeper A ring), and
//   * 8 KB instead of 12 KB of operand reads from shared memory per MMA step -- the K=256 GEMMs are bound by the
//     shared-memory port that operand fetch, TMA writes and the epilogue's TMA-store reads share.
// Pipeline (same roles as gemm_sm100.cuh): warp 0 = TMA producer in BOTH CTAs (each loads its own A slice; the
// transaction bytes of both land on the LEADER's full barrier), warp 1 lane 0 of the leader = MMA issuer
// (tcgen05.commit multicasts the "slot free" / "accumulator ready" arrivals to both CTAs), warps 2.. = epilogue in both
// CTAs on their own TMEM (the peer's epilogue warps release the accumulator stage with remote mbarrier arrives).
#pragma once
#include "gemm_sm100.cuh"
#include "ptx_cluster.cuh"

namespace svb {

template <uint32_t EPI_BYTES>
struct Gemm2Cfg {
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;      // one CTA's A slice of a k-block: 16 KB
  static constexpr uint32_t kBHalfBytes = 128 * kBlockK * 2;      // half of the 256-row weight tile per k-block: 16 KB
  static constexpr int kResidentKBlocks = 4;                      // K <= 256
  static constexpr uint32_t kResidentBytes = kResidentKBlocks * kBHalfBytes;
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kBarrierBytes = 256;
  static constexpr uint32_t kEpiBytes = (EPI_BYTES + 1023u) & ~1023u;
  static constexpr int kFit = static_cast<int>((kMaxDynSmem - kBarrierBytes - kEpiBytes - kResidentBytes) / kABytes);
  static constexpr int kStages = kFit < 8 ? kFit : 8;
  static constexpr uint32_t kSmemBytes = kResidentBytes + kStages * kABytes + kEpiBytes + kBarrierBytes;
  static_assert(kStages >= 4, "too little shared memory left for the A ring");
};

// grid = 2 * tiles_n * groups CTAs, cluster (2,1,1): pair = blockIdx.x / 2, N tile = pair % tiles_n, group = pair / tiles_n
template <bool B_MN, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + Epi::kWarps * 32, 1)
gemm2_bstat_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmProblem p,
                   const __grid_constant__ typename Epi::Params ep) {
  using Cfg = Gemm2Cfg<Epi::kSmemBytes>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int BLOCK_N = 256;
  static_assert((2 * STAGES + 5) * 8 + 8 <= Cfg::kBarrierBytes, "barrier region too small");

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_ring = smem + Cfg::kResidentBytes;
  uint8_t* epi_smem = smem_ring + STAGES * Cfg::kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + Cfg::kEpiBytes);  // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                                  // used in the leader only
  uint64_t* b_full_bar = tmem_empty_bar + 2;                                     // used in the leader only
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(b_full_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) / 32, 0);
  const int lane = static_cast<int>(threadIdx.x) % 32;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("svb: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // the leader's producer arrives once (expect_tx covers both CTAs' bytes)
      mbar_init(&empty_bar[s], 1);  // one multicast commit per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 2 * Epi::kWarps);  // epilogue warps of both CTAs
    }
    mbar_init(b_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem2_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tmem2_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barriers of both CTAs are initialised before anybody signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int pair = static_cast<int>(blockIdx.x) >> 1;
  const int fixed_n = pair % p.tiles_n;
  const int group = pair / p.tiles_n;
  const int groups = (static_cast<int>(gridDim.x) >> 1) / p.tiles_n;
  const int pair_tiles = (p.M + 2 * kBlockM - 1) / (2 * kBlockM);
  const int nkb = (p.K + kBlockK - 1) / kBlockK;
  auto tile_of = [&](int tp) -> TileInfo {
    TileInfo ti;
    ti.tile_n = fixed_n;
    ti.n0 = fixed_n * BLOCK_N;
    ti.m0 = tp * 2 * kBlockM + static_cast<int>(rank) * kBlockM;
    ti.tile_m = ti.m0 / kBlockM;
    ti.split = 0;
    ti.cta_slot = group * 2 + static_cast<int>(rank);
    return ti;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      const uint32_t bfull_leader = mapa_u32(smem_u32(b_full_bar), 0);
      if (group < pair_tiles) {  // this CTA's half of the weight tile, once: N rows [n0 + 128*rank, +128)
        if (leader) mbar_arrive_expect_tx(b_full_bar, 2u * nkb * Cfg::kBHalfBytes);
        const int nb0 = fixed_n * BLOCK_N + static_cast<int>(rank) * 128;
        for (int kb = 0; kb < nkb; ++kb) {
          uint8_t* sb = smem + kb * Cfg::kBHalfBytes;
          if constexpr (!B_MN) {
            tma2_load_2d(sb, &tmB, bfull_leader, kb * kBlockK, nb0);                      // box 64 k x 128 rows
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma2_load_2d(sb + j * 8192, &tmB, bfull_leader, nb0 + 64 * j, kb * kBlockK);
          }
        }
      }
      uint32_t stage = 0, phase = 0;
      long long w_empty = 0;
      (void)w_empty;
      for (int tp = group; tp < pair_tiles; tp += groups) {
        const TileInfo ti = tile_of(tp);
        for (int kb = 0; kb < nkb; ++kb) {
          SVB_TRACED_WAIT(w_empty, &empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem_ring + stage * Cfg::kABytes;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * Cfg::kABytes);
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (p.a_slab) tma2_load_3d(sa, &tmA, full_leader, 0, ti.m0, kb);
          else tma2_load_2d(sa, &tmA, full_leader, kb * kBlockK, ti.m0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
#ifdef SVB_GEMM_TRACE
      if (p.trace) p.trace[blockIdx.x * 4 + 0] = w_empty;
#endif
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread of the leader CTA)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kBlockM, BLOCK_N, false, B_MN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      long long w_full = 0, w_tmem = 0;
      (void)w_full; (void)w_tmem;
      if (group < pair_tiles) mbar_wait(b_full_bar, 0);
      for (int tp = group; tp < pair_tiles; tp += groups) {
        SVB_TRACED_WAIT(w_tmem, &tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < nkb; ++kb) {
          SVB_TRACED_WAIT(w_full, &full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem_ring + stage * Cfg::kABytes);
          const uint32_t b_base = smem_u32(smem + kb * Cfg::kBHalfBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(a_base + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_base + k * 2048, 8192, 1024)
                                        : make_smem_desc_sw128(b_base + k * 32, 16, 1024);
            umma2_f16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma2_commit_both(&empty_bar[stage]);                       // A slot reusable in both CTAs
          if (kb == nkb - 1) umma2_commit_both(&tmem_full_bar[acc]);  // accumulators complete in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
#ifdef SVB_GEMM_TRACE
      if (p.trace) { p.trace[blockIdx.x * 4 + 1] = w_full; p.trace[blockIdx.x * 4 + 2] = w_tmem; }
#endif
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs, own TMEM)
    constexpr int EW = Epi::kWarps;
    static_assert(EW == 8, "8 epilogue warps");
    constexpr int kChunksPerWarp = (BLOCK_N / 32) / (EW / 4);
    const int ew = warp - 2;
    const int wq = warp % 4;
    const int cgroup = ew / 4;
    const int row_in_tile = wq * 32 + lane;
    const int tid = ew * 32 + lane;
    Epi epi(ep, epi_smem, ew, BLOCK_N);
    uint32_t acc = 0, acc_phase = 0;
    long long w_acc = 0;
    (void)w_acc;
    uint32_t tmem_empty_leader[2];
    tmem_empty_leader[0] = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
    tmem_empty_leader[1] = mapa_u32(smem_u32(&tmem_empty_bar[1]), 0);
    if (Epi::kColVecs > 0 && group < pair_tiles) epi.colvec_fetch(p, tile_of(group), tid);
    for (int tp = group; tp < pair_tiles; tp += groups) {
      const TileInfo ti = tile_of(tp);
      if (Epi::kColVecs > 0) {
        epi.colvec_commit(acc, tid);
        epi_bar_sync(EW * 32);
        if (tp + groups < pair_tiles) epi.colvec_fetch(p, tile_of(tp + groups), tid);
      }
      SVB_TRACED_WAIT(w_acc, &tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int row = ti.m0 + row_in_tile;
      epi.begin_tile(p, ti, row, wq, lane);
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * BLOCK_N;
      if constexpr (epi_prefetches_acc<Epi>::value) {
        float v[2][32];
        const int c_begin = cgroup * kChunksPerWarp;
        if (ti.n0 + c_begin * 32 < p.N) tmem_ld_32x32(t_addr + c_begin * 32, v[0]);
#pragma unroll
        for (int ci = 0; ci < kChunksPerWarp; ++ci) {
          const int c = c_begin + ci;
          const int col0 = ti.n0 + c * 32;
          if (col0 < p.N) {
            tmem_ld_wait();
            if (ci + 1 < kChunksPerWarp && col0 + 32 < p.N) tmem_ld_32x32(t_addr + (c + 1) * 32, v[(ci + 1) & 1]);
            epi.chunk(p, ti, row, col0, v[ci & 1], wq, lane, ci);
          }
        }
      } else {
#pragma unroll
        for (int ci = 0; ci < kChunksPerWarp; ++ci) {
          const int c = cgroup * kChunksPerWarp + ci;
          const int col0 = ti.n0 + c * 32;
          if (col0 < p.N) {
            float v[32];
            if constexpr (!epi_skips_acc_load<Epi>::value) {
              tmem_ld_32x32(t_addr + c * 32, v);
              tmem_ld_wait();
            }
            epi.chunk(p, ti, row, col0, v, wq, lane, ci);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tmem_empty_leader[acc]);  // accumulator stage free: tell the leader's MMA thread
      epi.end_tile(p, ti, row, wq, lane);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    epi.finish(wq, lane);
#ifdef SVB_GEMM_TRACE
    if (p.trace && ew == 0 && lane == 0) p.trace[blockIdx.x * 4 + 3] = w_acc;
#endif
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still signal or issue MMAs on this CTA
  if (warp == 1) tmem2_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace svb

// ---------------------------------------------------------------------------------------------------------------- host
#include "gemm_host.cuh"
namespace svb {
// Two-CTA B-stationary launch (gemm2_sm100.cuh): K <= 256, A K-major (row-major pitch lda, or slab-major), B K-major
// [N, K] or MN-major [K, N] with pitch ldb.  *groups_out receives the number of pair groups per N tile (the epilogue's
// per-CTA partials have 2 * groups slots).  Returns 0 or a negative svb error code.
inline int gemm2_groups(int M, int N, int max_ctas = 0) {
  const int sms = max_ctas > 0 ? max_ctas : device_sm_count();
  const int tiles_n = (N + 255) / 256, pair_tiles = (M + 2 * kBlockM - 1) / (2 * kBlockM);
  if (tiles_n > sms / 2) return 0;
  int groups = (sms / 2) / tiles_n;
  if (groups > pair_tiles) groups = pair_tiles;
  return groups;
}
template <bool B_MN, class Epi>
int launch_gemm2_bstat(cudaStream_t stream, const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K,
                       const typename Epi::Params& ep, bool a_slab = false, int max_ctas = 0) {
  using Cfg = Gemm2Cfg<Epi::kSmemBytes>;
  if (M <= 0 || N <= 0 || K <= 0 || (N % 8) || K > Cfg::kResidentKBlocks * kBlockK) return -2;
  CUtensorMap tmA, tmB;
  int rc = a_slab ? make_tmap_bf16_slab(&tmA, A, M, K, kBlockM) : make_tmap_bf16_2d(&tmA, A, M, K, lda, kBlockM);
  if (rc) return rc;
  rc = !B_MN ? make_tmap_bf16_2d(&tmB, B, N, K, ldb, 128) : make_tmap_bf16_2d(&tmB, B, K, N, ldb, kBlockK);
  if (rc) return rc;
  GemmProblem p;
  p.M = M; p.N = N; p.K = K;
  p.k_splits = 1; p.k_per_split = ((K + kBlockK - 1) / kBlockK) * kBlockK;
  p.tiles_m = (M + kBlockM - 1) / kBlockM;
  p.tiles_n = (N + 255) / 256;
  p.a_slab = a_slab ? 1 : 0;
#ifdef SVB_GEMM_TRACE
  p.trace = gemm_trace_ptr();
#endif
  const int groups = gemm2_groups(M, N, max_ctas);
  if (groups < 1) return -5;
  auto kern = gemm2_bstat_kernel<B_MN, Epi>;
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return -4;
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) return -4;
    configured[dev] = true;
  }
  (kern<<<2 * p.tiles_n * groups, 64 + Epi::kWarps * 32, Cfg::kSmemBytes, stream>>>(tmA, tmB, p, ep), svb::count_launch());
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}
}  // namespace svb
