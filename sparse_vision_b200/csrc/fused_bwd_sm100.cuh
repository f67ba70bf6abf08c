// Fused backward of the SaeMLP encoder for C <= 256 (sm_100a):  dE GEMM -> ReLU mask -> dW_enc GEMM in ONE kernel, so
// that dPre' [T,F] (822 MB of writes + 822 MB of reads at cfg2) never exists in HBM (model_pipeline.py:385 autograd
// backward of models/sae_mlp.py:49-52).
//
// Transposed orientation: FEATURES are the accumulator rows (TMEM lanes).  A CTA owns one 128-feature tile and a strided
// set of 128-token blocks ("slot" s of S: blocks s, s+S, ...; the S * tiles_f CTAs that are resident together walk the
// same few token blocks, so DIFF / X tiles are read from HBM once and hit L2 for the other feature tiles).  Per block:
//   MMA1  acc1[128 f x 128 t] = W_dec^T[128 f x C] * DIFF^T[C x 128 t]          (A MN-major & resident, B K-major)
//   epilogue (8 warps): dPre' = mask(t,f) ? acc1 + l1c : 0  -> bf16 -> shared memory as a K-major / 128B-swizzled A tile;
//                       the per-feature sum over tokens (-> db_enc, rank-1 fix-up) stays inside the thread
//   MMA2  acc2[128 f x C] += P[128 f x 128 t] * X[128 t x C]                    (A K-major from smem, B MN-major)
// and at the end acc2 leaves as the split-K partial P_we[slot][f][c] (the format EpiPartial writes), colsum as
// colsum[2 * slot + token half][f].  The MMA thread issues MMA1(i+1) before MMA2(i): the tensor pipe works on the next
// block's dE tile while the epilogue warps turn block i into the P tile.
// TMEM: acc1 double-buffered (2 x 128 columns) + acc2 (256 columns).  Shared memory: W tile 64 KB resident, P 32 KB,
// DIFF 4 x 16 KB k-blocks, X 2 x 32 KB k-blocks (each with its own full / empty barrier, refilled one block ahead).
#pragma once
#include <cstdlib>
#include "gemm_host.cuh"
#include "epilogues.cuh"
#include "ptx_cluster.cuh"

namespace svb {

struct FusedBwdParams {
  int T, C, F;
  int slots;               // token-block slots per feature tile; grid = tiles_f * slots, every CTA gets >= 1 block
  int tiles_f;             // ceil(F / 128)
  int diff_slab, x_slab;   // operand stored slab-major (3-D tensor map)
  const uint32_t* mask;    // group-major 1-bit ReLU masks (mask_index)
  int words;               // ceil(F / 32)
  float l1c;
  float* part;             // [slots][F][C] fp32 split-K partials of dW_enc' = dPre'^T X
  long long part_stride;   // F * C
  float* colsum;           // [2 * slots][F]: sum over the CTA's tokens of dPre' (token halves separately)
  int prefetch;            // > 0: one CTA per slot pulls the DIFF / X tiles of the block this many steps ahead into L2
#ifdef SVB_FBW_TRACE
  long long* trace;        // bring-up only: [grid][8] cycles spent waiting (see tools/fused_bwd_probe.cu)
#endif
};
#ifdef SVB_FBW_TRACE
#define FBW_WAIT(slot_, call) do { const long long t0_ = clock64(); call; tr[slot_] += clock64() - t0_; } while (0)
#else
#define FBW_WAIT(slot_, call) call
#endif

namespace fbw {
constexpr uint32_t kWBytes = 4 * 16384;          // resident W_dec^T tile: 4 k-blocks of [2 atoms][64 c][64 f]
constexpr uint32_t kPBytes = 2 * 16384;          // P tile: 2 k-blocks of [128 f][64 t]
constexpr uint32_t kDBytes = 16384;              // one DIFF k-block [128 t][64 c]
constexpr uint32_t kXBytes = 32768;              // one X k-block: [C/64 atoms][64 t][64 c]
constexpr uint32_t kWOff = 0, kPOff = kWBytes, kDOff = kPOff + kPBytes, kXOff = kDOff + 4 * kDBytes,
                   kBarOff = kXOff + 2 * kXBytes;
constexpr uint32_t kSmem = kBarOff + 256;
static_assert(kSmem <= kMaxDynSmem, "fused backward: shared memory budget");
struct Bars {
  uint64_t w_full, d_full[4], d_empty[4], x_full[2], x_empty[2], acc1_full[2], acc1_empty[2], p_full, p_empty, acc2_full;
  uint32_t tmem_ptr;
};
static_assert(sizeof(Bars) <= 256, "barrier block");
}  // namespace fbw

__global__ void __launch_bounds__(320, 1)
fused_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmD,
                 const __grid_constant__ CUtensorMap tmX, const FusedBwdParams p) {
  using namespace fbw;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Ws = smem + kWOff;
  uint8_t* Ps = smem + kPOff;
  uint8_t* Ds = smem + kDOff;
  uint8_t* Xs = smem + kXOff;
  Bars* bar = reinterpret_cast<Bars*>(smem + kBarOff);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) / 32, 0);
  const int lane = static_cast<int>(threadIdx.x) % 32;

  const int ftile = static_cast<int>(blockIdx.x) % p.tiles_f;
  const int slot = static_cast<int>(blockIdx.x) / p.tiles_f;
  const int f0 = ftile * 128;
  const int nblocks = (p.T + 127) / 128;
  const int n = (nblocks - slot + p.slots - 1) / p.slots;   // token blocks of this CTA (>= 1 by construction)
  const int nkb = p.C / 64;                                   // C % 64 == 0, C <= 256

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("svb: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmD); tma_prefetch_desc(&tmX);
    mbar_init(&bar->w_full, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&bar->d_full[i], 1); mbar_init(&bar->d_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar->x_full[i], 1); mbar_init(&bar->x_empty[i], 1);
      mbar_init(&bar->acc1_full[i], 1); mbar_init(&bar->acc1_empty[i], 8);
    }
    mbar_init(&bar->p_full, 8); mbar_init(&bar->p_empty, 1);
    mbar_init(&bar->acc2_full, 1);
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
#ifdef SVB_FBW_TRACE
      long long tr[8] = {0};
#endif
      mbar_arrive_expect_tx(&bar->w_full, nkb * 16384u);
      for (int kb = 0; kb < nkb; ++kb)
        for (int j = 0; j < 2; ++j) tma_load_2d(Ws + kb * 16384 + j * 8192, &tmW, &bar->w_full, f0 + 64 * j, kb * 64);
      for (int i = 0; i <= n; ++i) {
        // The tiles_f CTAs of a slot walk the same token blocks at the same pace, and the first touch of a block is an
        // HBM read that every one of them would wait for: one of them (round robin) asks for it a few blocks ahead.
        if (p.prefetch > 0 && i + p.prefetch < n && (i + p.prefetch) % p.tiles_f == ftile) {
          const int t0 = (slot + (i + p.prefetch) * p.slots) * 128;
          for (int kb = 0; kb < nkb; ++kb) {
            if (p.diff_slab) tma_prefetch_l2_3d(&tmD, 0, t0, kb);
            else tma_prefetch_l2_2d(&tmD, kb * 64, t0);
          }
          for (int kb2 = 0; kb2 < 2; ++kb2)
            for (int j = 0; j < nkb; ++j) {
              if (p.x_slab) tma_prefetch_l2_3d(&tmX, 0, t0 + kb2 * 64, j);
              else tma_prefetch_l2_2d(&tmX, 64 * j, t0 + kb2 * 64);
            }
        }
        if (i < n) {   // DIFF tile of block i (B operand of MMA1, K-major)
          const int t0 = (slot + i * p.slots) * 128;
          const uint32_t par = static_cast<uint32_t>(i & 1) ^ 1u;
          for (int kb = 0; kb < nkb; ++kb) {
            FBW_WAIT(0, mbar_wait(&bar->d_empty[kb], par));
            mbar_arrive_expect_tx(&bar->d_full[kb], kDBytes);
            if (p.diff_slab) tma_load_3d(Ds + kb * kDBytes, &tmD, &bar->d_full[kb], 0, t0, kb);
            else tma_load_2d(Ds + kb * kDBytes, &tmD, &bar->d_full[kb], kb * 64, t0);
          }
        }
        if (i >= 1) {  // X tile of block i-1 (B operand of MMA2, MN-major: one 64 t x 64 c atom per 64 channels)
          const int t0 = (slot + (i - 1) * p.slots) * 128;
          const uint32_t par = static_cast<uint32_t>((i - 1) & 1) ^ 1u;
          for (int kb2 = 0; kb2 < 2; ++kb2) {
            FBW_WAIT(1, mbar_wait(&bar->x_empty[kb2], par));
            mbar_arrive_expect_tx(&bar->x_full[kb2], nkb * 8192u);
            for (int j = 0; j < nkb; ++j) {
              uint8_t* dst = Xs + kb2 * kXBytes + j * 8192;
              if (p.x_slab) tma_load_3d(dst, &tmX, &bar->x_full[kb2], 0, t0 + kb2 * 64, j);
              else tma_load_2d(dst, &tmX, &bar->x_full[kb2], 64 * j, t0 + kb2 * 64);
            }
          }
        }
      }
#ifdef SVB_FBW_TRACE
      if (p.trace) for (int q = 0; q < 2; ++q) p.trace[blockIdx.x * 8 + q] = tr[q];
#endif
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
#ifdef SVB_FBW_TRACE
      long long tr[8] = {0};
#endif
      constexpr uint32_t idesc1 = make_idesc_bf16(128, 128, true, false);
      const uint32_t idesc2 = make_idesc_bf16(128, p.C, false, true);
      const uint32_t acc2 = tmem_base + 256;
      mbar_wait(&bar->w_full, 0);
      for (int i = 0; i <= n; ++i) {
        if (i < n) {
          const int a = i & 1;
          FBW_WAIT(2, mbar_wait(&bar->acc1_empty[a], static_cast<uint32_t>((i >> 1) & 1) ^ 1u));
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + a * 128;
          for (int kb = 0; kb < nkb; ++kb) {
            FBW_WAIT(3, mbar_wait(&bar->d_full[kb], static_cast<uint32_t>(i & 1)));
            tc_fence_after();
            const uint32_t a_base = smem_u32(Ws + kb * 16384), b_base = smem_u32(Ds + kb * kDBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(d_tmem, make_smem_desc_sw128(a_base + k * 2048, 8192, 1024),
                       make_smem_desc_sw128(b_base + k * 32, 16, 1024), idesc1, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&bar->d_empty[kb]);
          }
          umma_commit(&bar->acc1_full[a]);
        }
        if (i >= 1) {
          const int ib = i - 1;
          FBW_WAIT(4, mbar_wait(&bar->p_full, static_cast<uint32_t>(ib & 1)));
          tc_fence_after();
          for (int kb2 = 0; kb2 < 2; ++kb2) {
            FBW_WAIT(5, mbar_wait(&bar->x_full[kb2], static_cast<uint32_t>(ib & 1)));
            tc_fence_after();
            const uint32_t a_base = smem_u32(Ps + kb2 * 16384), b_base = smem_u32(Xs + kb2 * kXBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(acc2, make_smem_desc_sw128(a_base + k * 32, 16, 1024),
                       make_smem_desc_sw128(b_base + k * 2048, 8192, 1024), idesc2, (ib | kb2 | k) != 0 ? 1u : 0u);
            umma_commit(&bar->x_empty[kb2]);
          }
          umma_commit(&bar->p_empty);
        }
      }
      umma_commit(&bar->acc2_full);
#ifdef SVB_FBW_TRACE
      if (p.trace) for (int q = 2; q < 6; ++q) p.trace[blockIdx.x * 8 + q] = tr[q];
#endif
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 4 lane quarters x 2 token halves
    const int ew = warp - 2, wq = warp % 4, h = ew / 4;
    const int r = wq * 32 + lane;                        // feature row inside the tile
    const int f = f0 + r;
    const int w = (f0 >> 5) + wq;                        // mask word of this warp's 32 features
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    float csum = 0.f;
#ifdef SVB_FBW_TRACE
    long long tr[8] = {0};
#endif
    // mask words of this warp's 64 tokens: lane l holds the words of tokens l and 32 + l (fetched one block ahead)
    auto load_words = [&](int i, uint32_t (&mw)[2]) {
      const long long t0 = static_cast<long long>(slot + i * p.slots) * 128 + h * 64;
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const long long t = t0 + ci * 32 + lane;
        mw[ci] = (t < p.T && w < p.words) ? __ldg(p.mask + mask_index(t, w, p.T)) : 0u;
      }
    };
    uint32_t mw_next[2];
    load_words(0, mw_next);
    for (int i = 0; i < n; ++i) {
      uint32_t mw[2] = {mw_next[0], mw_next[1]};
      if (i + 1 < n) load_words(i + 1, mw_next);
      const int a = i & 1;
      FBW_WAIT(6, mbar_wait(&bar->acc1_full[a], static_cast<uint32_t>((i >> 1) & 1)));
      tc_fence_after();
      float v[2][32];
      tmem_ld_32x32(lane_base + a * 128 + h * 64, v[0]);
      tmem_ld_32x32(lane_base + a * 128 + h * 64 + 32, v[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar->acc1_empty[a]);   // the MMA warp may overwrite this accumulator stage
      uint32_t pk[2][16];
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        // lane l holds the mask word of token l (bit f = feature f of this warp): transposed, this lane's word has
        // bit j = token j of its own feature
        const uint32_t tw = transpose_bits32(mw[ci], lane);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = (tw & (1u << j)) ? v[ci][j] + p.l1c : 0.f;
          v[ci][j] = x;
          csum += x;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) pk[ci][q] = pack_bf16x2(v[ci][2 * q], v[ci][2 * q + 1]);
      }
      if (i >= 1) FBW_WAIT(7, mbar_wait(&bar->p_empty, static_cast<uint32_t>((i - 1) & 1)));   // MMA2 of the previous block has read P
      uint8_t* row = Ps + h * 16384 + r * 128;
#pragma unroll
      for (int ci = 0; ci < 2; ++ci)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(row + (((ci * 4 + q) ^ (r & 7)) << 4)) =
              make_uint4(pk[ci][4 * q], pk[ci][4 * q + 1], pk[ci][4 * q + 2], pk[ci][4 * q + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar->p_full);
    }
    // the CTA's split-K partial of dW_enc' and its column sums
    mbar_wait(&bar->acc2_full, 0);
    tc_fence_after();
#ifdef SVB_FBW_TRACE
    if (p.trace && ew == 0 && lane == 0) for (int q = 6; q < 8; ++q) p.trace[blockIdx.x * 8 + q] = tr[q];
#endif
    if (f < p.F) p.colsum[static_cast<size_t>(2 * slot + h) * p.F + f] = csum;
    const int nchunks = p.C / 32;
    for (int c = h; c < nchunks; c += 2) {
      float v[32];
      tmem_ld_32x32(lane_base + 256 + c * 32, v);
      tmem_ld_wait();
      if (f < p.F) store_row_f32(p.part + slot * p.part_stride + static_cast<long long>(f) * p.C + c * 32, v, 32);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ two-CTA variant
// cta_group::2 version for C % 128 == 0: an SM pair owns 256 features (CTA r: features f0 + 128 r ...).  Both MMAs run
// with M = 256, so each CTA holds only HALF of every streamed B operand -- 64 of the block's 128 tokens of DIFF for
// MMA1 (N = 128 tokens), C/2 of the channels of X for MMA2 (N = C) -- which halves the L2 -> SM traffic (the single-CTA
// kernel pulls 128 KB per 2 x 1024 MMA cycles, ~9.5 TB/s chip-wide at cfg2, against an L2 limit of ~12 TB/s) and lets
// the same 128 KB of streaming shared memory hold TWO token blocks instead of one.  The leader's MMA thread issues for
// the pair; "full" barriers live in the leader (both producers credit them), "empty" / "accumulator ready" arrivals are
// multicast commits, and the epilogue warps of both CTAs release acc1 / publish their P tiles with cluster-scope
// arrives on the leader's barriers.
namespace fbw2 {
constexpr uint32_t kWOff = 0, kPOff = 65536, kDOff = kPOff + 32768, kXOff = kDOff + 2 * 32768, kBarOff = kXOff + 2 * 32768;
constexpr uint32_t kSmem = kBarOff + 256;
static_assert(kSmem <= kMaxDynSmem, "fused backward (2 CTA): shared memory budget");
struct Bars {
  uint64_t w_full, d_full[2], d_empty[2], x_full[2], x_empty[2], acc1_full[2], acc1_empty[2], p_full, p_empty, acc2_full;
  uint32_t tmem_ptr;
};
static_assert(sizeof(Bars) <= 256, "barrier block");
}  // namespace fbw2

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
fused_bwd2_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmD,
                  const __grid_constant__ CUtensorMap tmX, const FusedBwdParams p) {
  using namespace fbw2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* Ws = smem + kWOff;
  uint8_t* Ps = smem + kPOff;
  uint8_t* Ds = smem + kDOff;     // 2 stages x [nkb k-blocks][64 t][64 c]
  uint8_t* Xs = smem + kXOff;     // 2 stages x [2 k-blocks of 64 t][C/128 atoms][64 t][64 c]
  Bars* bar = reinterpret_cast<Bars*>(smem + kBarOff);
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) / 32, 0);
  const int lane = static_cast<int>(threadIdx.x) % 32;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  const int pair = static_cast<int>(blockIdx.x) >> 1;
  const int ptile = pair % p.tiles_f;                  // tiles_f counts 256-feature pair tiles here
  const int slot = pair / p.tiles_f;
  const int f0 = ptile * 256 + static_cast<int>(rank) * 128;
  const int nblocks = (p.T + 127) / 128;
  const int n = (nblocks - slot + p.slots - 1) / p.slots;
  const int nkb = p.C / 64;                            // C % 128 == 0, C <= 256
  const int nat = p.C / 128;                           // 64-channel atoms of X in this CTA's half

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("svb: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmD); tma_prefetch_desc(&tmX);
    mbar_init(&bar->w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar->d_full[i], 1); mbar_init(&bar->d_empty[i], 1);
      mbar_init(&bar->x_full[i], 1); mbar_init(&bar->x_empty[i], 1);
      mbar_init(&bar->acc1_full[i], 1); mbar_init(&bar->acc1_empty[i], 16);   // epilogue warps of both CTAs
    }
    mbar_init(&bar->p_full, 16); mbar_init(&bar->p_empty, 1);
    mbar_init(&bar->acc2_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem2_alloc(&bar->tmem_ptr, 512); tmem2_relinquish(); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // barriers of both CTAs are initialised before anybody signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = bar->tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs, own halves)
    if (lane == 0) {
#ifdef SVB_FBW_TRACE
      long long tr[8] = {0};
#endif
      const uint32_t w_full_l = mapa_u32(smem_u32(&bar->w_full), 0);
      if (leader) mbar_arrive_expect_tx(&bar->w_full, 2u * nkb * 16384u);
      for (int kb = 0; kb < nkb; ++kb)
        for (int j = 0; j < 2; ++j) tma2_load_2d(Ws + kb * 16384 + j * 8192, &tmW, w_full_l, f0 + 64 * j, kb * 64);
      for (int i = 0; i <= n; ++i) {
        if (i < n) {   // this CTA's 64 tokens of the DIFF tile of block i
          const int s = i & 1;
          const int t0 = (slot + i * p.slots) * 128 + static_cast<int>(rank) * 64;
          FBW_WAIT(0, mbar_wait(&bar->d_empty[s], static_cast<uint32_t>((i >> 1) & 1) ^ 1u));
          if (leader) mbar_arrive_expect_tx(&bar->d_full[s], 2u * nkb * 8192u);
          const uint32_t full_l = mapa_u32(smem_u32(&bar->d_full[s]), 0);
          for (int kb = 0; kb < nkb; ++kb) {
            uint8_t* dst = Ds + s * 32768 + kb * 8192;
            if (p.diff_slab) tma2_load_3d(dst, &tmD, full_l, 0, t0, kb);
            else tma2_load_2d(dst, &tmD, full_l, kb * 64, t0);
          }
        }
        if (i >= 1) {  // this CTA's C/2 channels of the X tile of block i-1
          const int ib = i - 1, s = ib & 1;
          const int t0 = (slot + ib * p.slots) * 128;
          FBW_WAIT(1, mbar_wait(&bar->x_empty[s], static_cast<uint32_t>((ib >> 1) & 1) ^ 1u));
          if (leader) mbar_arrive_expect_tx(&bar->x_full[s], 2u * 2u * nat * 8192u);
          const uint32_t full_l = mapa_u32(smem_u32(&bar->x_full[s]), 0);
          for (int kb2 = 0; kb2 < 2; ++kb2)
            for (int j = 0; j < nat; ++j) {
              uint8_t* dst = Xs + s * 32768 + kb2 * 16384 + j * 8192;
              const int ja = static_cast<int>(rank) * nat + j;   // 64-channel atom of the whole X tile
              if (p.x_slab) tma2_load_3d(dst, &tmX, full_l, 0, t0 + kb2 * 64, ja);
              else tma2_load_2d(dst, &tmX, full_l, 64 * ja, t0 + kb2 * 64);
            }
        }
      }
#ifdef SVB_FBW_TRACE
      if (p.trace) for (int q = 0; q < 2; ++q) p.trace[blockIdx.x * 8 + q] = tr[q];
#endif
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread of the leader CTA)
    if (leader && lane == 0) {
#ifdef SVB_FBW_TRACE
      long long tr[8] = {0};
#endif
      constexpr uint32_t idesc1 = make_idesc_bf16(256, 128, true, false);
      const uint32_t idesc2 = make_idesc_bf16(256, p.C, false, true);
      const uint32_t acc2 = tmem_base + 256;
      mbar_wait(&bar->w_full, 0);
      for (int i = 0; i <= n; ++i) {
        if (i < n) {
          const int a = i & 1;
          const uint32_t par = static_cast<uint32_t>((i >> 1) & 1);
          FBW_WAIT(2, mbar_wait(&bar->acc1_empty[a], par ^ 1u));
          FBW_WAIT(3, mbar_wait(&bar->d_full[a], par));
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + a * 128;
          for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t a_base = smem_u32(Ws + kb * 16384), b_base = smem_u32(Ds + a * 32768 + kb * 8192);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma2_f16(d_tmem, make_smem_desc_sw128(a_base + k * 2048, 8192, 1024),
                        make_smem_desc_sw128(b_base + k * 32, 16, 1024), idesc1, (kb | k) != 0 ? 1u : 0u);
          }
          umma2_commit_both(&bar->d_empty[a]);
          umma2_commit_both(&bar->acc1_full[a]);
        }
        if (i >= 1) {
          const int ib = i - 1, s = ib & 1;
          FBW_WAIT(4, mbar_wait(&bar->p_full, static_cast<uint32_t>(ib & 1)));   // P tiles of BOTH CTAs are in place
          FBW_WAIT(5, mbar_wait(&bar->x_full[s], static_cast<uint32_t>((ib >> 1) & 1)));
          tc_fence_after();
          for (int kb2 = 0; kb2 < 2; ++kb2) {
            const uint32_t a_base = smem_u32(Ps + kb2 * 16384), b_base = smem_u32(Xs + s * 32768 + kb2 * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma2_f16(acc2, make_smem_desc_sw128(a_base + k * 32, 16, 1024),
                        make_smem_desc_sw128(b_base + k * 2048, 8192, 1024), idesc2, (ib | kb2 | k) != 0 ? 1u : 0u);
          }
          umma2_commit_both(&bar->x_empty[s]);
          umma2_commit_both(&bar->p_empty);
        }
      }
      umma2_commit_both(&bar->acc2_full);
#ifdef SVB_FBW_TRACE
      if (p.trace) for (int q = 2; q < 6; ++q) p.trace[blockIdx.x * 8 + q] = tr[q];
#endif
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs, own TMEM)
    const int ew = warp - 2, wq = warp % 4, h = ew / 4;
    const int r = wq * 32 + lane;
    const int f = f0 + r;
    const int w = (f0 >> 5) + wq;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t acc1_empty_l[2] = {mapa_u32(smem_u32(&bar->acc1_empty[0]), 0), mapa_u32(smem_u32(&bar->acc1_empty[1]), 0)};
    const uint32_t p_full_l = mapa_u32(smem_u32(&bar->p_full), 0);
    float csum = 0.f;
#ifdef SVB_FBW_TRACE
    long long tr[8] = {0};
#endif
    auto load_words = [&](int i, uint32_t (&mw)[2]) {
      const long long t0 = static_cast<long long>(slot + i * p.slots) * 128 + h * 64;
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const long long t = t0 + ci * 32 + lane;
        mw[ci] = (t < p.T && w < p.words) ? __ldg(p.mask + mask_index(t, w, p.T)) : 0u;
      }
    };
    uint32_t mw_next[2];
    load_words(0, mw_next);
    for (int i = 0; i < n; ++i) {
      uint32_t mw[2] = {mw_next[0], mw_next[1]};
      if (i + 1 < n) load_words(i + 1, mw_next);
      const int a = i & 1;
      FBW_WAIT(6, mbar_wait(&bar->acc1_full[a], static_cast<uint32_t>((i >> 1) & 1)));
      tc_fence_after();
      float v[2][32];
      tmem_ld_32x32(lane_base + a * 128 + h * 64, v[0]);
      tmem_ld_32x32(lane_base + a * 128 + h * 64 + 32, v[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc1_empty_l[a]);
      uint32_t pk[2][16];
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        // lane l holds the mask word of token l (bit f = feature f of this warp): transposed, this lane's word has
        // bit j = token j of its own feature
        const uint32_t tw = transpose_bits32(mw[ci], lane);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = (tw & (1u << j)) ? v[ci][j] + p.l1c : 0.f;
          v[ci][j] = x;
          csum += x;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) pk[ci][q] = pack_bf16x2(v[ci][2 * q], v[ci][2 * q + 1]);
      }
      if (i >= 1) FBW_WAIT(7, mbar_wait(&bar->p_empty, static_cast<uint32_t>((i - 1) & 1)));
      uint8_t* row = Ps + h * 16384 + r * 128;
#pragma unroll
      for (int ci = 0; ci < 2; ++ci)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(row + (((ci * 4 + q) ^ (r & 7)) << 4)) =
              make_uint4(pk[ci][4 * q], pk[ci][4 * q + 1], pk[ci][4 * q + 2], pk[ci][4 * q + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(p_full_l);
    }
    mbar_wait(&bar->acc2_full, 0);
    tc_fence_after();
#ifdef SVB_FBW_TRACE
    if (p.trace && ew == 0 && lane == 0) for (int q = 6; q < 8; ++q) p.trace[blockIdx.x * 8 + q] = tr[q];
#endif
    if (f < p.F) p.colsum[static_cast<size_t>(2 * slot + h) * p.F + f] = csum;
    const int nchunks = p.C / 32;
    for (int c = h; c < nchunks; c += 2) {
      float v[32];
      tmem_ld_32x32(lane_base + 256 + c * 32, v);
      tmem_ld_wait();
      if (f < p.F) store_row_f32(p.part + slot * p.part_stride + static_cast<long long>(f) * p.C + c * 32, v, 32);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // nobody leaves (or frees TMEM) while the peer may still signal or issue MMAs on this CTA
  if (warp == 1) tmem2_dealloc(tmem_base, 512);
}

// tuning(kTuneFusedBwd) = 0 keeps the two un-fused GEMMs (A/B measurements, tests).
inline bool fused_bwd_enabled() { return tuning(kTuneFusedBwd) != 0; }

#ifdef SVB_FBW_TRACE
inline long long*& fused_bwd_trace_ptr() {
  static long long* p = nullptr;
  return p;
}
#endif
// measured: no gain inside the step (0.369-0.376 ms at 0..6 blocks ahead), so 0 by default
inline int fused_bwd_prefetch() { return tuning(kTuneFbwPrefetch); }

// tuning(kTuneFbwTwoCta) = 0 keeps the single-CTA kernel for every shape.
inline bool fused_bwd_two_cta(int C) { return tuning(kTuneFbwTwoCta) != 0 && C % 128 == 0; }
// Slots the launcher will use (sizes the split-K workspaces): S = min(sms / tiles_f, token blocks), or with SM pairs
// and 256-feature pair tiles for the two-CTA kernel.
inline int fused_bwd_slots(long long T, int C, int F, int max_ctas = 0) {
  const int sms = max_ctas > 0 ? max_ctas : device_sm_count();
  const bool two = fused_bwd_two_cta(C);
  const int units = two ? sms / 2 : sms;
  const int tiles_f = two ? (F + 255) / 256 : (F + 127) / 128;
  if (tiles_f > units) return 0;
  long long s = units / tiles_f;
  const long long nblocks = (T + 127) / 128;
  return static_cast<int>(s < nblocks ? s : nblocks);
}
inline bool fused_bwd_supported(long long T, int C, int F, int max_ctas = 0) {
  return C % 64 == 0 && C >= 64 && C <= 256 && F % 8 == 0 && T > 0 && T < (1ll << 31) - 256 && fused_bwd_slots(T, C, F, max_ctas) >= 1;
}

// W_dec bf16 [C, F] row-major; DIFF [T, C] and X [T, C] bf16 row-major (pitch ldd / ldx) or slab-major.
// part: [slots][F][C] fp32, colsum: [2 * slots][F].  Returns 0 or a negative code.
inline int launch_fused_bwd(cudaStream_t stream, const void* w_dec, const void* diff, bool diff_slab, int64_t ldd,
                            const void* x, bool x_slab, int64_t ldx, const uint32_t* mask, int T, int C, int F, float l1c,
                            float* part, float* colsum, int max_ctas = 0) {
  if (!fused_bwd_supported(T, C, F, max_ctas)) return -2;
  const bool two = fused_bwd_two_cta(C);
  const uint32_t d_rows = two ? 64 : 128;   // token rows of one DIFF box: each CTA of a pair loads half of the block
  CUtensorMap tmW, tmD, tmX;
  int rc = make_tmap_bf16_2d(&tmW, w_dec, C, F, F, 64);
  if (rc) return rc;
  rc = diff_slab ? make_tmap_bf16_slab(&tmD, diff, T, C, d_rows) : make_tmap_bf16_2d(&tmD, diff, T, C, ldd, d_rows);
  if (rc) return rc;
  rc = x_slab ? make_tmap_bf16_slab(&tmX, x, T, C, 64) : make_tmap_bf16_2d(&tmX, x, T, C, ldx, 64);
  if (rc) return rc;
  FusedBwdParams p;
  p.T = T; p.C = C; p.F = F;
  p.tiles_f = two ? (F + 255) / 256 : (F + 127) / 128;
  p.slots = fused_bwd_slots(T, C, F, max_ctas);
  p.diff_slab = diff_slab ? 1 : 0; p.x_slab = x_slab ? 1 : 0;
  p.mask = mask; p.words = (F + 31) / 32; p.l1c = l1c;
  p.prefetch = fused_bwd_prefetch();
#ifdef SVB_FBW_TRACE
  p.trace = fused_bwd_trace_ptr();
#endif
  p.part = part; p.part_stride = static_cast<long long>(F) * C; p.colsum = colsum;
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return -4;
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(fused_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fbw::kSmem) != cudaSuccess ||
        cudaFuncSetAttribute(fused_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fbw2::kSmem) != cudaSuccess)
      return -4;
    configured[dev] = true;
  }
  if (two) (fused_bwd2_kernel<<<2 * p.tiles_f * p.slots, 320, fbw2::kSmem, stream>>>(tmW, tmD, tmX, p), svb::count_launch());
  else (fused_bwd_kernel<<<p.tiles_f * p.slots, 320, fbw::kSmem, stream>>>(tmW, tmD, tmX, p), svb::count_launch());
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // namespace svb
