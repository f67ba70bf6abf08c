// Fused SaeMLP forward for C <= 256 (sm_100a):  encoder GEMM -> bias / ReLU / mask -> decoder GEMM in ONE two-CTA
// kernel, so that E [T,F] is written once and never re-read by a decoder pass (models/sae_mlp.py:49-52).
//
// An SM pair (cta_group::2) owns 256 tokens (CTA r: tokens m0 + 128 r ...) and keeps their X rows resident (<= 64 KB per
// CTA).  It walks the 256-feature tiles j of F:
//   GEMM1(j)  acc1[256 t x 256 f] = X * W_enc[j]^T                       (TMEM columns 0..255 of both CTAs)
//   epilogue  e = relu(acc1 + fold) -> 1-bit mask words, sum|e|, bf16 e written ONCE into shared memory in the UMMA
//             K-major / 128B-swizzle layout; from there it is TMA-stored to the slab-major E workspace (the tile IS a
//             stack of slab pieces) and consumed as the A operand of
//   GEMM2(j)  acc2[256 t x C] += E_tile * W_dec[:, j]^T                  (TMEM columns 256..511)
// and after the last feature tile the decoder epilogue (EpiDecNchw, unchanged: + b_dec, DIFF = d - x, statistics,
// NCHW / channels_last / channel-major output) runs once per token tile on acc2.  The MMA thread issues
// GEMM1(j) -> GEMM2(j-1) -> GEMM1(j+1) ...: the tensor pipe works on the previous tile's decoder contribution while the
// epilogue warps read acc1, and acc1 is released as soon as its 4 chunks per warp are in registers.
// Weights are streamed from L2 through a ring of 16 KB stages; with M = 256 each CTA loads only HALF of every weight
// tile (128 of the 256 feature rows of W_enc[j], C/2 of the channel rows of W_dec[:, j]): 128 KB per CTA and feature
// tile instead of 256 KB -- the un-paired bring-up version (tools/fused_fwd_probe.cu, round 1) was bound by exactly
// that stream.
//
// STATUS: bit-exact against naive kernels on five shapes (tools/fused_fwd_probe.cu), but NOT on the product path: at cfg2 it
// takes 0.44 ms against 0.41 ms for the separate encoder and decoder GEMMs.  TMEM holds acc1 (256 columns) and ONE acc2
// (256 columns), so the decoder epilogue of a token tile (17 kcycles) cannot overlap the next tile's GEMM2s, and the
// encoder-style epilogue (2.8 kcycles per 128 x 256 tile with 8 warps, no faster with 16) leaves the tensor pipe at 58 %
// even with every store stubbed out (0.32 ms).  DESIGN.md section 8 has the measurements.
#pragma once
#include <cstdlib>
#include "../sparse_vision_b200/csrc/gemm_host.cuh"
#include "../sparse_vision_b200/csrc/epilogues.cuh"
#include "../sparse_vision_b200/csrc/ptx_cluster.cuh"

namespace svb {

struct FusedFwdParams {
  int T, C, F;
  int pair_tiles;          // ceil(T / 256)
  int nf;                  // F / 256
  int x_slab;              // X stored slab-major (3-D tensor map)
  const float* fold;       // [F] folded encoder bias  b_enc - W_enc b_dec
  uint32_t* mask;          // group-major 1-bit ReLU masks (mask_index), rows = T
  int words;               // F / 32
  float* l1_partial;       // [gridDim.x * 16]: sum |e| per CTA and epilogue warp
#ifdef SVB_FFW_TRACE
  long long* trace;        // bring-up only: [grid][16] cycles spent waiting (see tools/fused_fwd_probe.cu)
  int dbg;                 // bring-up only (timing, wrong results): 1 = no E store, 2 = no decoder epilogue work, 4 = no mask store
#endif
};
#ifdef SVB_FFW_TRACE
#define FFW_WAIT(slot_, call) do { const long long t0_ = clock64(); call; tr[slot_] += clock64() - t0_; } while (0)
#define FFW_DBG(bit) (p.dbg & (bit))
#define FFW_TRACE_DECL long long tr[16] = {0}; const long long tr_start = clock64(); (void)tr_start
#define FFW_TRACE_OUT(lo, hi) do { if (p.trace) for (int q_ = lo; q_ < hi; ++q_) p.trace[blockIdx.x * 16 + q_] = tr[q_]; } while (0)
#else
#define FFW_WAIT(slot_, call) call
#define FFW_DBG(bit) false
#define FFW_TRACE_DECL
#define FFW_TRACE_OUT(lo, hi)
#endif

namespace ffw {
constexpr int kEpiWarps = 16;                    // 4 per TMEM lane quarter: each owns 64 of a tile's 256 columns
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;
constexpr int kStages = 6;
constexpr uint32_t kStage = 16384;
constexpr uint32_t kXOff = 0, kEOff = 65536, kRingOff = 131072, kCvOff = kRingOff + kStages * kStage, kBarOff = kCvOff + 2048;
constexpr uint32_t kSmem = kBarOff + 256;
static_assert(kSmem <= kMaxDynSmem, "fused forward: shared memory budget");
using DecEpi = EpiDecNchwT<kEpiWarps>;
static_assert(kEpiWarps * 4096 <= 65536, "the decoder epilogue's staging aliases the E tile");
struct Bars {
  uint64_t x_full, x_empty, full[kStages], empty[kStages], acc1_full, acc1_empty, es_full, es_empty, acc2_full, acc2_empty;
  uint32_t tmem_ptr;
};
static_assert(sizeof(Bars) <= 256, "barrier block");
}  // namespace ffw

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ffw::kThreads, 1)
fused_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWe,
                 const __grid_constant__ CUtensorMap tmWd, const __grid_constant__ CUtensorMap tmE,
                 const FusedFwdParams p, const __grid_constant__ EpiDecNchwParams dp) {
  using namespace ffw;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Xs = smem + kXOff;      // [C/64 k-blocks][128 t][64 c]
  uint8_t* Es = smem + kEOff;      // [4 k-blocks][128 t][64 f]; between token tiles: the decoder epilogue's staging
  uint8_t* ring = smem + kRingOff;
  float* cv_enc = reinterpret_cast<float*>(smem + kCvOff);   // 2 x 256 floats: -fold of the current / next feature tile
  Bars* bar = reinterpret_cast<Bars*>(smem + kBarOff);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) / 32, 0);
  const int lane = static_cast<int>(threadIdx.x) % 32;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = static_cast<int>(blockIdx.x) >> 1, npairs = static_cast<int>(gridDim.x) >> 1;
  const int nkb = p.C / 64;
  const int NF = p.nf;
  const uint32_t wd_bytes = static_cast<uint32_t>(p.C / 2) * 128u;   // one W_dec half k-block: C/2 channel rows x 64 features

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("svb: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmWe); tma_prefetch_desc(&tmWd); tma_prefetch_desc(&tmE);
    tma_prefetch_desc(&dp.tm_diff); tma_prefetch_desc(&dp.tm_out);
    mbar_init(&bar->x_full, 1); mbar_init(&bar->x_empty, 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(&bar->full[s], 1); mbar_init(&bar->empty[s], 1); }
    mbar_init(&bar->acc1_full, 1); mbar_init(&bar->acc1_empty, 2 * kEpiWarps);   // epilogue warps of both CTAs
    mbar_init(&bar->es_full, 2 * kEpiWarps); mbar_init(&bar->es_empty, 1);
    mbar_init(&bar->acc2_full, 1); mbar_init(&bar->acc2_empty, 2 * kEpiWarps);
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
      uint32_t stage = 0, phase = 0;
      FFW_TRACE_DECL;
      const uint32_t x_full_l = mapa_u32(smem_u32(&bar->x_full), 0);
      auto load_w = [&](const CUtensorMap* tm, int col, int row, uint32_t bytes) {
        FFW_WAIT(0, mbar_wait(&bar->empty[stage], phase ^ 1));
        if (leader) mbar_arrive_expect_tx(&bar->full[stage], 2u * bytes);
        tma2_load_2d(ring + stage * kStage, tm, mapa_u32(smem_u32(&bar->full[stage]), 0), col, row);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      };
      int it = 0;
      for (int pt = pair; pt < p.pair_tiles; pt += npairs, ++it) {
        const int m0 = pt * 256 + static_cast<int>(rank) * 128;
        FFW_WAIT(1, mbar_wait(&bar->x_empty, static_cast<uint32_t>(it & 1) ^ 1u));
        if (leader) mbar_arrive_expect_tx(&bar->x_full, 2u * nkb * 16384u);
        for (int kb = 0; kb < nkb; ++kb) {
          if (p.x_slab) tma2_load_3d(Xs + kb * 16384, &tmX, x_full_l, 0, m0, kb);
          else tma2_load_2d(Xs + kb * 16384, &tmX, x_full_l, kb * 64, m0);
        }
        for (int j = 0; j <= NF; ++j) {
          if (j < NF)   // W_enc [F, C]: this CTA's 128 feature rows of tile j
            for (int kb = 0; kb < nkb; ++kb) load_w(&tmWe, kb * 64, j * 256 + static_cast<int>(rank) * 128, 16384u);
          if (j >= 1)   // W_dec [C, F]: this CTA's C/2 channel rows, the 256 feature columns of tile j-1
            for (int kb2 = 0; kb2 < 4; ++kb2) load_w(&tmWd, (j - 1) * 256 + kb2 * 64, static_cast<int>(rank) * (p.C / 2), wd_bytes);
        }
      }
      FFW_TRACE_OUT(0, 2);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread of the leader CTA)
    if (leader && lane == 0) {
      const uint32_t idesc1 = make_idesc_bf16(256, 256, false, false);
      const uint32_t idesc2 = make_idesc_bf16(256, p.C, false, false);
      const uint32_t acc1 = tmem_base, acc2 = tmem_base + 256;
      uint32_t stage = 0, phase = 0;
      FFW_TRACE_DECL;
      auto mma_blocks = [&](uint32_t d_tmem, const uint8_t* a_tile, int nblk, uint32_t idesc, bool first_accumulates) {
        for (int kb = 0; kb < nblk; ++kb) {
          FFW_WAIT(3, mbar_wait(&bar->full[stage], phase));
          tc_fence_after();
          const uint32_t a_base = smem_u32(a_tile + kb * 16384), b_base = smem_u32(ring + stage * kStage);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma2_f16(d_tmem, make_smem_desc_sw128(a_base + k * 32, 16, 1024), make_smem_desc_sw128(b_base + k * 32, 16, 1024),
                      idesc, (first_accumulates || (kb | k) != 0) ? 1u : 0u);
          umma2_commit_both(&bar->empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      };
      int it = 0;
      for (int pt = pair; pt < p.pair_tiles; pt += npairs, ++it) {
        auto gemm2 = [&](int jj) {
          const int gg = it * NF + jj;
          FFW_WAIT(4, mbar_wait(&bar->es_full, static_cast<uint32_t>(gg & 1)));
          if (jj == 0) FFW_WAIT(5, mbar_wait(&bar->acc2_empty, static_cast<uint32_t>(it & 1) ^ 1u));
          tc_fence_after();
          mma_blocks(acc2, Es, 4, idesc2, jj != 0);
          umma2_commit_both(&bar->es_empty);
        };
        mbar_wait(&bar->x_full, static_cast<uint32_t>(it & 1));
        for (int j = 0; j < NF; ++j) {
          const int g = it * NF + j;
          FFW_WAIT(2, mbar_wait(&bar->acc1_empty, static_cast<uint32_t>(g & 1) ^ 1u));
          tc_fence_after();
          mma_blocks(acc1, Xs, nkb, idesc1, false);
          umma2_commit_both(&bar->acc1_full);
          if (j == NF - 1) umma2_commit_both(&bar->x_empty);
          if (j >= 1) gemm2(j - 1);
        }
        gemm2(NF - 1);
        umma2_commit_both(&bar->acc2_full);
      }
      FFW_TRACE_OUT(2, 6);
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 4 lane quarters x 4 column groups, both CTAs
    constexpr int kCh = 8 / (kEpiWarps / 4);             // 32-column chunks per warp and 256-column tile (2)
    const int ew = warp - 2, wq = warp % 4, cgroup = ew / 4;
    const int r = wq * 32 + lane;                        // token row inside this CTA's tile
    const int tid = ew * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t acc1_empty_l = mapa_u32(smem_u32(&bar->acc1_empty), 0);
    const uint32_t es_full_l = mapa_u32(smem_u32(&bar->es_full), 0);
    const uint32_t acc2_empty_l = mapa_u32(smem_u32(&bar->acc2_empty), 0);
    GemmProblem g2{};   // what the decoder epilogue sees: the decoder GEMM D[T, C]
    g2.M = p.T; g2.N = p.C; g2.K = p.F;
    DecEpi depi(dp, Es, ew, 256);
    {   // decoder bias: fetched once into registers, committed to shared memory before every decoder epilogue
      TileInfo t0{};
      depi.colvec_fetch(g2, t0, tid);
    }
    float l1_total = 0.f;
    FFW_TRACE_DECL;
    float nfold_next = tid < 256 ? 0.f - __ldg(p.fold + tid) : 0.f;   // -fold of feature tile 0
    int it = 0;
    for (int pt = pair; pt < p.pair_tiles; pt += npairs, ++it) {
      const int m0 = pt * 256 + static_cast<int>(rank) * 128;
      const int row = m0 + r;
      const bool row_ok = row < p.T;
      for (int j = 0; j < NF; ++j) {
        const int g = it * NF + j;
        // -fold of this feature tile into shared memory (fetched during the previous tile), next tile's on its way
        float* cv = cv_enc + (g & 1) * 256;
        if (tid < 256) cv[tid] = nfold_next;
        FFW_WAIT(8, epi_bar_sync(kEpiThreads));
        if (tid < 256) {
          const int jn = j + 1 < NF ? j + 1 : 0;   // wraps to tile 0 of the next token tile (same vector)
          // 0 - x: a zero bias stages +0, whose difference with acc = 0 is +0 (mask bit 0)
          nfold_next = 0.f - __ldg(p.fold + jn * 256 + tid);
        }
        FFW_WAIT(6, mbar_wait(&bar->acc1_full, static_cast<uint32_t>(g & 1)));
        tc_fence_after();
#ifdef SVB_FFW_TRACE
        const long long ta0 = clock64();
#endif
        uint32_t pk[kCh][16];
        uint32_t words[kCh];
        float sum = 0.f;
#pragma unroll
        for (int ci = 0; ci < kCh; ++ci) {
          float v[32], nb[32];
          tmem_ld_32x32(lane_base + (cgroup * kCh + ci) * 32, v);
          lds_row_f32(cv + (cgroup * kCh + ci) * 32, nb);
          tmem_ld_wait();
          // t = (-fold) - acc = -pre: its sign bit is set exactly when pre > 0; e = max(-t, 0)   (as EpiEncT)
          uint32_t wq4[4] = {0, 0, 0, 0};
          float sq4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int jj = 7; jj >= 0; --jj) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int i = q * 8 + jj;
              const float t = nb[i] - v[i];
              wq4[q] = __funnelshift_l(__float_as_uint(t), wq4[q], 1);
              v[i] = fmaxf(-t, 0.f);
              sq4[q] += v[i];
            }
          }
          words[ci] = row_ok ? ((wq4[0] | (wq4[1] << 8)) | ((wq4[2] << 16) | (wq4[3] << 24))) : 0u;
          sum += (sq4[0] + sq4[1]) + (sq4[2] + sq4[3]);
#pragma unroll
          for (int q = 0; q < 16; ++q) pk[ci][q] = pack_bf16x2(v[2 * q], v[2 * q + 1]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc1_empty_l);   // GEMM1 of the next feature tile may overwrite acc1
        if (row_ok && !FFW_DBG(4)) {
          l1_total += sum;
          static_assert(kCh == 2, "mask words are stored as one 8-byte pair per row and warp");
          *reinterpret_cast<uint2*>(p.mask + mask_index(row, j * 8 + cgroup * 2, p.T)) = make_uint2(words[0], words[1]);
        }
#ifdef SVB_FFW_TRACE
        tr[12] += clock64() - ta0;
#endif
        // the E tile is free once GEMM2 of the previous feature tile has read it and this warp's own TMA store has
        FFW_WAIT(7, mbar_wait(&bar->es_empty, static_cast<uint32_t>(g & 1) ^ 1u));
        FFW_WAIT(8, if (lane == 0) bulk_wait_read<0>(); __syncwarp());
#ifdef SVB_FFW_TRACE
        const long long ts0 = clock64();
#endif
        uint8_t* rowp = Es + cgroup * 16384 + r * 128;       // this warp's 64 columns are k-block `cgroup` of the tile
#pragma unroll
        for (int ci = 0; ci < kCh; ++ci)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint4*>(rowp + (((ci * 4 + i) ^ (r & 7)) << 4)) =
                make_uint4(pk[ci][4 * i], pk[ci][4 * i + 1], pk[ci][4 * i + 2], pk[ci][4 * i + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(es_full_l);                  // GEMM2 may read the tile (all epilogue warps of the pair arrive)
          if (!FFW_DBG(1)) tma_store_3d(&tmE, Es + cgroup * 16384 + wq * 4096, 0, m0 + wq * 32, j * 4 + cgroup);   // own 32 x 64 slab piece
          bulk_commit();
        }
#ifdef SVB_FFW_TRACE
        tr[13] += clock64() - ts0;
#endif
      }
      // ---- decoder epilogue of this token tile (acc2 complete: every GEMM2 has finished reading the E tile)
      FFW_WAIT(9, mbar_wait(&bar->acc2_full, static_cast<uint32_t>(it & 1)));
      tc_fence_after();
#ifdef SVB_FFW_TRACE
      const long long td0 = clock64();
#endif
      if (lane == 0) bulk_wait_read<0>();                  // ... and so has this warp's E store
      __syncwarp();
      // the bias vector goes into the -fold buffer that the last encoder step used (the next step writes the other one)
      depi.use_colvec_at(cv_enc + ((it * NF + NF - 1) & 1) * 256);
      epi_bar_sync(kEpiThreads);                           // every warp's E stores are done: the staging area is free,
      depi.colvec_commit(0, tid);                          // and nobody reads -fold any more
      epi_bar_sync(kEpiThreads);
      TileInfo ti{};
      ti.m0 = m0; ti.n0 = 0; ti.tile_m = m0 / 128; ti.tile_n = 0; ti.split = 0; ti.cta_slot = 0;
      const int nchunks = p.C / 32;
      if (m0 < p.T && !FFW_DBG(2)) {
        for (int c = cgroup; c < nchunks; c += kEpiWarps / 4) {
          float v[32];
          tmem_ld_32x32(lane_base + 256 + c * 32, v);
          tmem_ld_wait();
          depi.chunk(g2, ti, row, c * 32, v, wq, lane, 0);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(acc2_empty_l);
        bulk_wait_read<0>();                               // staging reads done before the next tile's E writes
      }
      __syncwarp();
      epi_bar_sync(kEpiThreads);
#ifdef SVB_FFW_TRACE
      tr[10] += clock64() - td0;
#endif
    }
#ifdef SVB_FFW_TRACE
    tr[11] = clock64() - tr_start;
    if (ew == 0 && lane == 0) FFW_TRACE_OUT(6, 14);
#endif
    depi.finish(wq, lane);
    if (p.l1_partial) {
      const float s = warp_sum(l1_total);
      if (lane == 0) p.l1_partial[static_cast<size_t>(blockIdx.x) * kEpiWarps + ew] = s;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem2_dealloc(tmem_base, 512);
}

#ifdef SVB_FFW_TRACE
inline long long*& fused_fwd_trace_ptr() {
  static long long* p = nullptr;
  return p;
}
#endif
#ifdef SVB_FFW_TRACE
inline int& fused_fwd_dbg() {
  static int v = 0;
  return v;
}
#endif
inline bool fused_fwd_supported(long long T, int C, int F, int max_ctas = 0) {
  const int sms = max_ctas > 0 ? max_ctas : device_sm_count();
  return C % 64 == 0 && C >= 64 && C <= 256 && F % 256 == 0 && F >= 256 && T > 0 && T < (1ll << 31) - 512 && sms >= 2;
}

// X [T, C] bf16 (row-major pitch ldx, or slab-major), W_enc bf16 [F, C], W_dec bf16 [C, F], E slab-major [F/64][T][64]
// (out).  dp: the decoder epilogue's parameters exactly as for the un-fused decoder GEMM, except that sq_partial (like
// l1_partial) has 16 entries per CTA.  Returns 0 or a negative code.
inline int launch_fused_fwd(cudaStream_t stream, const void* x, bool x_slab, int64_t ldx, const void* w_enc, const void* w_dec,
                            void* e_slab, const float* fold, uint32_t* mask, float* l1_partial, int T, int C, int F,
                            const EpiDecNchwParams& dp, int max_ctas = 0) {
  if (!fused_fwd_supported(T, C, F, max_ctas)) return -2;
  CUtensorMap tmX, tmWe, tmWd, tmE;
  int rc = x_slab ? make_tmap_bf16_slab(&tmX, x, T, C, 128) : make_tmap_bf16_2d(&tmX, x, T, C, ldx, 128);
  if (rc) return rc;
  if ((rc = make_tmap_bf16_2d(&tmWe, w_enc, F, C, C, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmWd, w_dec, C, F, F, C / 2))) return rc;
  if ((rc = make_tmap_bf16_slab(&tmE, e_slab, T, F, 32))) return rc;
  FusedFwdParams p;
  p.T = T; p.C = C; p.F = F;
  p.pair_tiles = (T + 255) / 256;
  p.nf = F / 256;
  p.x_slab = x_slab ? 1 : 0;
  p.fold = fold; p.mask = mask; p.words = F / 32; p.l1_partial = l1_partial;
#ifdef SVB_FFW_TRACE
  p.trace = fused_fwd_trace_ptr();
  p.dbg = fused_fwd_dbg();
#endif
  const int sms = max_ctas > 0 ? max_ctas : device_sm_count();
  int pairs = sms / 2;
  if (pairs > p.pair_tiles) pairs = p.pair_tiles;
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return -4;
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ffw::kSmem) != cudaSuccess) return -4;
    configured[dev] = true;
  }
  (fused_fwd_kernel<<<2 * pairs, ffw::kThreads, ffw::kSmem, stream>>>(tmX, tmWe, tmWd, tmE, p, dp), svb::count_launch());
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace svb
