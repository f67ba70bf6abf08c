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
#include <cstdlib>
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


// ------------------------------------------------------------------------------------------------ streaming variant
// Two-CTA version of the STREAMING schedule of gemm_bf16_kernel (any K, optional split-K):  an SM pair computes a
// 256 x 256 output tile (CTA r: rows m0 + 128 r ...), both operands come through the ring, and each CTA loads its own
// 128 rows of A but only HALF of the B tile (128 of the 256 N rows / columns).  Why: with K > 256 every single-CTA tile
// re-streams a full 256-wide B tile from L2 -- the decoder GEMM pulls 1.5 MB per 128-token tile (11.6 TB/s chip-wide at
// cfg2), the weight-gradient GEMMs 48 KB per k-block -- and L2 -> SM bandwidth (~6300 B/clk) is what bounds them.  Pairs
// cut that by a third (32 instead of 48 KB per k-block and CTA) and the smaller stages deepen the ring from 4 to 6.
template <uint32_t EPI_BYTES>
struct Gemm2StreamCfg {
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;   // this CTA's 128 rows of A: 16 KB
  static constexpr uint32_t kBBytes = 128 * kBlockK * 2;       // this CTA's half of the B tile: 16 KB
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kBarrierBytes = 256;
  static constexpr uint32_t kEpiBytes = (EPI_BYTES + 1023u) & ~1023u;
  static constexpr int kFit = static_cast<int>((kMaxDynSmem - kBarrierBytes - kEpiBytes) / kStageBytes);
  static constexpr int kStages = kFit < 6 ? kFit : 6;
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarrierBytes;
  static_assert(kStages >= 3, "too little shared memory left for the operand ring");
};

// p.tiles_m counts 256-row PAIR tiles here.  grid = 2 * pairs, cluster (2,1,1).
template <bool A_MN, bool B_MN, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + Epi::kWarps * 32, 1)
gemm2_stream_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmProblem p,
                    const __grid_constant__ typename Epi::Params ep) {
  using Cfg = Gemm2StreamCfg<Epi::kSmemBytes>;
  constexpr int STAGES = Cfg::kStages;
  constexpr int BLOCK_N = 256;
  static_assert((2 * STAGES + 4) * 8 + 8 <= Cfg::kBarrierBytes, "barrier region too small");
  static_assert(!epi_ones_col<Epi>::value, "the ones-column epilogues run on the single-CTA kernel");

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* epi_smem = smem + STAGES * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + Cfg::kEpiBytes);  // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                                  // used in the leader only
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

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
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 2 * Epi::kWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem2_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tmem2_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int pair = static_cast<int>(blockIdx.x) >> 1, npairs = static_cast<int>(gridDim.x) >> 1;
  const int num_tiles = p.tiles_m * p.tiles_n * p.k_splits;
  auto decode = [&](int t) -> TileInfo {
    TileInfo ti;
    ti.tile_n = t % p.tiles_n;
    const int r = t / p.tiles_n;
    const int tp = p.reverse_m ? p.tiles_m - 1 - r % p.tiles_m : r % p.tiles_m;
    ti.split = r / p.tiles_m;
    ti.m0 = tp * 2 * kBlockM + static_cast<int>(rank) * kBlockM;
    ti.tile_m = ti.m0 / kBlockM;
    ti.n0 = ti.tile_n * BLOCK_N;
    ti.cta_slot = 0;
    return ti;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int t = pair; t < num_tiles; t += npairs) {
        const TileInfo ti = decode(t);
        const int k_begin = ti.split * p.k_per_split;
        const int k_end = min(p.K, k_begin + p.k_per_split);
        const int nkb = (k_end - k_begin + kBlockK - 1) / kBlockK;
        const int nb0 = ti.n0 + static_cast<int>(rank) * 128;   // this CTA's half of the B tile
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * Cfg::kStageBytes);
          const uint32_t full_l = mapa_u32(smem_u32(&full_bar[stage]), 0);
          const int k0 = k_begin + kb * kBlockK;
          if constexpr (!A_MN) {
            if (p.a_slab) tma2_load_3d(sa, &tmA, full_l, 0, ti.m0, k0 >> 6);
            else tma2_load_2d(sa, &tmA, full_l, k0, ti.m0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (p.a_slab) tma2_load_3d(sa + j * 8192, &tmA, full_l, 0, k0, (ti.m0 >> 6) + j);
              else tma2_load_2d(sa + j * 8192, &tmA, full_l, ti.m0 + 64 * j, k0);
            }
          }
          if constexpr (!B_MN) {
            if (p.b_slab) tma2_load_3d(sb, &tmB, full_l, 0, nb0, k0 >> 6);
            else tma2_load_2d(sb, &tmB, full_l, k0, nb0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              if (p.b_slab) tma2_load_3d(sb + j * 8192, &tmB, full_l, 0, k0, (nb0 >> 6) + j);
              else tma2_load_2d(sb + j * 8192, &tmB, full_l, nb0 + 64 * j, k0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread of the leader CTA)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kBlockM, BLOCK_N, A_MN, B_MN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int t = pair; t < num_tiles; t += npairs) {
        const TileInfo ti = decode(t);
        const int k_begin = ti.split * p.k_per_split;
        const int k_end = min(p.K, k_begin + p.k_per_split);
        const int nkb = (k_end - k_begin + kBlockK - 1) / kBlockK;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_base = a_base + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t adesc = A_MN ? make_smem_desc_sw128(a_base + k * 2048, 8192, 1024)
                                        : make_smem_desc_sw128(a_base + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_base + k * 2048, 8192, 1024)
                                        : make_smem_desc_sw128(b_base + k * 32, 16, 1024);
            umma2_f16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma2_commit_both(&empty_bar[stage]);
          if (kb == nkb - 1) umma2_commit_both(&tmem_full_bar[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
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
    const int n_lim = epi_pads_n64<Epi>::value ? ((p.N + 63) & ~63) : p.N;
    uint32_t acc = 0, acc_phase = 0;
    uint32_t tmem_empty_leader[2];
    tmem_empty_leader[0] = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0);
    tmem_empty_leader[1] = mapa_u32(smem_u32(&tmem_empty_bar[1]), 0);
    if (Epi::kColVecs > 0 && pair < num_tiles) epi.colvec_fetch(p, decode(pair), tid);
    for (int t = pair; t < num_tiles; t += npairs) {
      const TileInfo ti = decode(t);
      if (Epi::kColVecs > 0) {
        epi.colvec_commit(acc, tid);
        epi_bar_sync(EW * 32);
        if (t + npairs < num_tiles) epi.colvec_fetch(p, decode(t + npairs), tid);
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int row = ti.m0 + row_in_tile;
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * BLOCK_N;
      // a tile that lies entirely below the matrix (second CTA of the last pair) has nothing to store or reduce
      const bool live = ti.m0 < p.M;
      if (live) epi.begin_tile(p, ti, row, wq, lane);
      if constexpr (epi_prefetches_acc<Epi>::value) {
        float v[2][32];
        const int c_begin = cgroup * kChunksPerWarp;
        if (live && ti.n0 + c_begin * 32 < n_lim) tmem_ld_32x32(t_addr + c_begin * 32, v[0]);
#pragma unroll
        for (int ci = 0; ci < kChunksPerWarp; ++ci) {
          const int c = c_begin + ci;
          const int col0 = ti.n0 + c * 32;
          if (live && col0 < n_lim) {
            tmem_ld_wait();
            if (ci + 1 < kChunksPerWarp && col0 + 32 < n_lim) tmem_ld_32x32(t_addr + (c + 1) * 32, v[(ci + 1) & 1]);
            epi.chunk(p, ti, row, col0, v[ci & 1], wq, lane, ci);
          }
        }
      } else {
#pragma unroll
        for (int ci = 0; ci < kChunksPerWarp; ++ci) {
          const int c = cgroup * kChunksPerWarp + ci;
          const int col0 = ti.n0 + c * 32;
          if (live && col0 < n_lim) {
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
      if (lane == 0) mbar_arrive_cluster(tmem_empty_leader[acc]);
      if (live) epi.end_tile(p, ti, row, wq, lane);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    epi.finish(wq, lane);
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
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

// Streaming two-CTA launch: same operand conventions as launch_gemm (A K-major [M, K] / MN-major [K, M], B likewise with
// N; row-major pitches lda / ldb or slab-major).  Split count as planned_splits2() reports it.
inline int planned_splits2(int M, int N, int K, int k_splits_req, int max_ctas = 0) {
  const int tiles_mp = (M + 2 * kBlockM - 1) / (2 * kBlockM), tiles_n = (N + 255) / 256;
  const int kblocks = (K + kBlockK - 1) / kBlockK;
  const int pairs = (max_ctas > 0 ? max_ctas : device_sm_count()) / 2;
  int splits = k_splits_req;
  if (splits <= 0) {
    const int mn_tiles = tiles_mp * tiles_n;
    splits = mn_tiles >= pairs ? 1 : (pairs / mn_tiles);
  }
  if (splits > kblocks) splits = kblocks;
  if (splits < 1) splits = 1;
  const int kb_per = (kblocks + splits - 1) / splits;
  return (kblocks + kb_per - 1) / kb_per;
}
template <bool A_MN, bool B_MN, class Epi>
int launch_gemm2_stream(cudaStream_t stream, const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K,
                        int k_splits_req, const typename Epi::Params& ep, int* splits_out = nullptr, int max_ctas = 0,
                        bool a_slab = false, bool b_slab = false, bool reverse_m = false) {
  using Cfg = Gemm2StreamCfg<Epi::kSmemBytes>;
  if (M <= 0 || N <= 0 || K <= 0 || (N % 8)) return -2;
  CUtensorMap tmA, tmB;
  int rc;
  if (a_slab) rc = !A_MN ? make_tmap_bf16_slab(&tmA, A, M, K, kBlockM) : make_tmap_bf16_slab(&tmA, A, K, M, kBlockK);
  else if (!A_MN) rc = make_tmap_bf16_2d(&tmA, A, M, K, lda, kBlockM);
  else rc = make_tmap_bf16_2d(&tmA, A, K, M, lda, kBlockK);
  if (rc) return rc;
  if (b_slab) rc = !B_MN ? make_tmap_bf16_slab(&tmB, B, N, K, 128) : make_tmap_bf16_slab(&tmB, B, K, N, kBlockK);
  else if (!B_MN) rc = make_tmap_bf16_2d(&tmB, B, N, K, ldb, 128);
  else rc = make_tmap_bf16_2d(&tmB, B, K, N, ldb, kBlockK);
  if (rc) return rc;
  GemmProblem p;
  p.M = M; p.N = N; p.K = K;
  p.a_slab = a_slab ? 1 : 0;
  p.b_slab = b_slab ? 1 : 0;
  p.reverse_m = reverse_m ? 1 : 0;
  p.tiles_m = (M + 2 * kBlockM - 1) / (2 * kBlockM);   // pair tiles
  p.tiles_n = (N + 255) / 256;
  const int kblocks = (K + kBlockK - 1) / kBlockK;
  const int splits = planned_splits2(M, N, K, k_splits_req, max_ctas);
  const int kb_per = (kblocks + splits - 1) / splits;
  p.k_splits = splits;
  p.k_per_split = kb_per * kBlockK;
  if (splits_out) *splits_out = splits;
  auto kern = gemm2_stream_kernel<A_MN, B_MN, Epi>;
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return -4;
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) return -4;
    configured[dev] = true;
  }
  const int pairs_max = (max_ctas > 0 ? max_ctas : device_sm_count()) / 2;
  if (pairs_max < 1) return -5;
  const int num_tiles = p.tiles_m * p.tiles_n * p.k_splits;
  const int pairs = num_tiles < pairs_max ? num_tiles : pairs_max;
  (kern<<<2 * pairs, 64 + Epi::kWarps * 32, Cfg::kSmemBytes, stream>>>(tmA, tmB, p, ep), svb::count_launch());
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// tuning(kTuneGemmPairs) = 0 keeps every streaming GEMM on the single-CTA kernel.
inline bool gemm2_stream_enabled() { return tuning(kTuneGemmPairs) != 0; }
// The streaming GEMMs of the training steps go through here: SM pairs when the problem has more than one 128-row tile
// (with M <= 128 the second CTA of every pair would idle), the single-CTA kernel otherwise.  Same argument list as
// launch_gemm<256, ...>.
inline bool gemm_uses_pairs(int M, int max_ctas = 0) {
  return gemm2_stream_enabled() && M > kBlockM && (max_ctas > 0 ? max_ctas : device_sm_count()) >= 2;
}
inline int planned_splits_s(int M, int N, int K, int k_splits_req, int max_ctas = 0) {
  return gemm_uses_pairs(M, max_ctas) ? planned_splits2(M, N, K, k_splits_req, max_ctas) : planned_splits<256>(M, N, K, k_splits_req, max_ctas);
}
template <bool A_MN, bool B_MN, class Epi>
int launch_gemm_s(cudaStream_t stream, const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K,
                  int k_splits_req, const typename Epi::Params& ep, int* splits_out = nullptr, int max_ctas = 0,
                  unsigned long long a_policy = 0, bool a_slab = false, bool b_slab = false, int a_prefetch = 0,
                  bool reverse_m = false) {
  if (gemm_uses_pairs(M, max_ctas))
    return launch_gemm2_stream<A_MN, B_MN, Epi>(stream, A, lda, B, ldb, M, N, K, k_splits_req, ep, splits_out, max_ctas, a_slab,
                                                 b_slab, reverse_m);
  return launch_gemm<256, A_MN, B_MN, Epi>(stream, A, lda, B, ldb, M, N, K, k_splits_req, ep, splits_out, max_ctas, a_policy,
                                           a_slab, b_slab, a_prefetch, reverse_m);
}
}  // namespace svb
