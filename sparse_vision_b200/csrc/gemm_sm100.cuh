// Persistent, warp-specialised bf16 GEMM for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T  (fp32 accumulate in TMEM).
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : MMA issuer    (one thread issues tcgen05.mma, tcgen05.commit frees smem slots / publishes TMEM)
//   warps 2..    : epilogue     (tcgen05.ld TMEM -> registers -> Epi functor -> global).  Epi::kWarps = 8 or 16
//                 warps (2 or 4 per SM sub-partition): warps that share a TMEM lane quarter split the tile's columns.
//                 A K=256 tile gives the epilogue only 2048 MMA cycles, and the per-warp latency chain (TMEM load,
//                 bias, ReLU, mask, sums, bf16 pack, slab store) of 4 chunks is ~3x that; 16 warps with 2 chunks each
//                 and per-column vectors staged one tile ahead bring it under the MMA time.
//
// Either operand may be K-major (K contiguous in memory) or MN-major (M/N contiguous), which covers every GEMM of
// the SAE step on row-major token tensors without a transpose copy:
//   enc   pre = X   [T,C] * W_enc[F,C]^T      A K-major,  B K-major
//   dec   d   = E   [T,F] * W_dec[C,F]^T      A K-major,  B K-major
//   dE        = dD  [T,C] * W_dec[C,F]        A K-major,  B MN-major
//   dW_dec    = dD^T[C,T] * E   [T,F]         A MN-major, B MN-major   (split-K over tokens)
//   dW_enc    = dP^T[F,T] * X   [T,C]         A MN-major, B MN-major   (split-K over tokens)
// The accumulator is double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps the MMAs of
// tile i+1.  The epilogue is a functor: see epilogues.cuh.
#pragma once
#include "ptx.cuh"

namespace svb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;

struct GemmProblem {
  int M, N, K;
  int k_splits;     // >= 1; every split is non-empty
  int k_per_split;  // multiple of kBlockK
  int tiles_m, tiles_n;
  unsigned long long a_policy = 0, b_policy = 0;  // L2 eviction policy of the operand loads (0 = default)
  int a_slab = 0, b_slab = 0;  // operand stored slab-major ([cols/64][rows][64], 3-D tensor map; see gemm_host.cuh)
  int a_prefetch = 0;          // B-stationary schedule: L2-prefetch the A tiles this many steps of the CTA's walk ahead
  int reverse_m = 0;           // streaming schedule: walk the M tiles from the last to the first (the kernel that wrote
                               // A just before left its LAST tiles in L2)
#ifdef SVB_GEMM_TRACE
  long long* trace = nullptr;  // bring-up only: [gridDim.x][4] cycles the producer / MMA thread / epilogue spent waiting
  int a_skip = 0;              // bring-up only (timing, wrong results): load only every a_skip-th A stage from memory
#endif
};
#ifdef SVB_GEMM_TRACE
#define SVB_TRACED_WAIT(acc, bar, par) do { const long long t0_ = clock64(); mbar_wait(bar, par); acc += clock64() - t0_; } while (0)
#else
#define SVB_TRACED_WAIT(acc, bar, par) mbar_wait(bar, par)
#endif

struct TileInfo {
  int m0, n0;     // element offsets of this tile
  int split;      // split-K slice
  int tile_m;     // tile index along M
  int tile_n;
  int cta_slot;   // B-stationary schedules: index of this CTA among the CTAs that share its N tile
};

constexpr uint32_t kMaxDynSmem = 232448;  // 227 KB per CTA on sm_100

// BSTAT ("B-stationary"): for K <= 256 the whole [BLOCK_N x K] B tile (<= 128 KB) stays resident in shared memory
// and a CTA walks the M tiles of ONE N tile, so only A is streamed.  Shared-memory bandwidth (~128 B/clk/SM shared
// by TMA writes, UMMA operand reads and the epilogue's staging) is what bounds the K=256 GEMMs; not re-writing B
// for every tile removes a quarter of that traffic.
template <int BLOCK_N, uint32_t EPI_BYTES, bool BSTAT = false>
struct GemmCfg {
  static constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
  static constexpr uint32_t kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kResidentKBlocks = BSTAT ? 4 : 0;  // K <= 256
  static constexpr uint32_t kResidentBytes = kResidentKBlocks * kBBytes;
  static constexpr uint32_t kStageBytes = BSTAT ? kABytes : kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 2 * BLOCK_N;
  static constexpr uint32_t kBarrierBytes = 256;  // 2*stages + 5 mbarriers + tmem ptr
  static constexpr uint32_t kEpiBytes = (EPI_BYTES + 1023u) & ~1023u;  // epilogue staging, 1024-aligned (TMA swizzle)
  static constexpr int kFit = static_cast<int>((kMaxDynSmem - kBarrierBytes - kEpiBytes - kResidentBytes) / kStageBytes);
  static constexpr int kCap = BSTAT ? 8 : ((BLOCK_N == 256) ? 4 : 6);
  static constexpr int kStages = kFit < kCap ? kFit : kCap;
  static constexpr uint32_t kSmemBytes = kResidentBytes + kStages * kStageBytes + kEpiBytes + kBarrierBytes;
  static_assert(kStages >= 2, "epilogue staging leaves no room for a pipelined operand ring");
};

// Named barrier among the epilogue threads only (id 1; id 0 is __syncthreads).
__device__ __forceinline__ void epi_bar_sync(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// Per-column float vectors (biases ...) of a tile, staged through registers into shared memory ONE TILE AHEAD:
// fetch() issues the global loads for the next tile right after the barrier of the current one, commit() writes them
// to smem at the top of the next iteration, so the L2 round trip (there is next to no L1 beside ~200 KB of dynamic
// smem) overlaps a whole tile of epilogue work.  dst alternates between two buffers by accumulator stage; the one
// barrier per tile also bounds the skew between epilogue warps to less than a tile, which makes that safe.
template <int NV, int NTHREADS>
struct ColVecStage {
  static constexpr int kPer = (NV * 256 + NTHREADS - 1) / NTHREADS;
  float r[kPer];
  __device__ __forceinline__ void fetch(const float* const (&src)[NV], int n0, int N, int tid) {
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int idx = tid + i * NTHREADS;
      const int v = idx >> 8, c = idx & 255;
      r[i] = (idx < NV * 256 && src[v < NV ? v : 0] != nullptr && n0 + c < N) ? __ldg(src[v < NV ? v : 0] + n0 + c) : 0.f;
    }
  }
  __device__ __forceinline__ void commit(float* dst, int tid) const {
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int idx = tid + i * NTHREADS;
      if (idx < NV * 256) dst[idx] = r[i];
    }
  }
};

// Epilogues may define `static constexpr bool kSkipAccLoad = true` (bring-up probes only) to skip the TMEM read.
template <class E, class = void> struct epi_skips_acc_load { static constexpr bool value = false; };
template <class E> struct epi_skips_acc_load<E, decltype(void(E::kSkipAccLoad))> { static constexpr bool value = E::kSkipAccLoad; };
// Epilogues may define `static constexpr bool kOnesCol = true`: besides D = A B^T the kernel accumulates the ROW SUMS
// of A over K (A times a column of ones, one extra N = 16 MMA per k-step on a constant shared-memory tile) into TMEM
// column BLOCK_N and hands them to Epi::row_sum().  The accumulator is then single-buffered (the second half of TMEM
// holds the extra column), which is free for the split-K weight-gradient GEMMs: every CTA computes one tile.
template <class E, class = void> struct epi_ones_col { static constexpr bool value = false; };
template <class E> struct epi_ones_col<E, decltype(void(E::kOnesCol))> { static constexpr bool value = E::kOnesCol; };
// Epilogues may define `static constexpr bool kPrefetchAcc = true` to double-buffer the TMEM reads in registers.
// Epilogues with `static constexpr bool kPadN64 = true` write a slab-major output and are also handed the all-zero
// accumulator chunks between N and the end of N's last 64-column slab, so that the slab's padding columns get defined
// (zero) contents: the matrix is the K operand of a later GEMM.
template <class E, class = void> struct epi_pads_n64 { static constexpr bool value = false; };
template <class E> struct epi_pads_n64<E, decltype(void(E::kPadN64))> { static constexpr bool value = E::kPadN64; };
template <class E, class = void> struct epi_prefetches_acc { static constexpr bool value = false; };
template <class E> struct epi_prefetches_acc<E, decltype(void(E::kPrefetchAcc))> { static constexpr bool value = E::kPrefetchAcc; };

__device__ __forceinline__ TileInfo decode_tile(const GemmProblem& p, int t, int block_n) {
  TileInfo ti;
  ti.tile_n = t % p.tiles_n;
  const int r = t / p.tiles_n;
  ti.tile_m = p.reverse_m ? p.tiles_m - 1 - r % p.tiles_m : r % p.tiles_m;
  ti.split = r / p.tiles_m;
  ti.m0 = ti.tile_m * kBlockM;
  ti.n0 = ti.tile_n * block_n;
  ti.cta_slot = 0;
  return ti;
}

template <int BLOCK_N, bool A_MN, bool B_MN, class Epi, bool BSTAT = false>
__global__ void __launch_bounds__(64 + Epi::kWarps * 32, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmProblem p, const __grid_constant__ typename Epi::Params ep) {
  using Cfg = GemmCfg<BLOCK_N, Epi::kSmemBytes, BSTAT>;
  constexpr int STAGES = Cfg::kStages;
  static_assert(BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N must be 128 or 256");
  static_assert((2 * STAGES + 5) * 8 + 8 <= Cfg::kBarrierBytes, "barrier region too small");

  // No static __shared__ anywhere in this kernel, so the dynamic window starts at the CTA's shared-memory base and
  // the 1024-byte alignment the 128B swizzle needs holds without slack (checked once below).
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_ring = smem + Cfg::kResidentBytes;              // operand ring (after the resident B tile, if any)
  uint8_t* epi_smem = smem_ring + STAGES * Cfg::kStageBytes;    // 1024-aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + Cfg::kEpiBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* b_full_bar = tmem_empty_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(b_full_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x) / 32, 0);
  const int lane = static_cast<int>(threadIdx.x) % 32;

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
      mbar_init(&tmem_empty_bar[a], Epi::kWarps);
    }
    mbar_init(b_full_bar, 1);
    fence_barrier_init();
  }
  constexpr bool ONES = epi_ones_col<Epi>::value;
  if constexpr (ONES) {  // 16 rows x 128 B of bf16 1.0 at the start of the epilogue staging area (any swizzle of ones is ones)
    static_assert(Epi::kSmemBytes >= 2048, "the ones tile lives in the epilogue staging area");
    if (warp >= 2 && threadIdx.x - 64 < 128) {
      reinterpret_cast<uint4*>(epi_smem)[threadIdx.x - 64] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
      fence_proxy_async_smem();
    }
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_tiles = p.tiles_m * p.tiles_n * p.k_splits;
  // Tile sequence of this CTA.  Streaming: t = blockIdx.x, += gridDim.x over all (split, m, n) tiles.
  // B-stationary: the CTA owns N tile blockIdx.x % tiles_n and walks M tiles t = blockIdx.x / tiles_n, += groups.
  const int fixed_n = BSTAT ? static_cast<int>(blockIdx.x) % p.tiles_n : 0;
  const int t_first = BSTAT ? static_cast<int>(blockIdx.x) / p.tiles_n : static_cast<int>(blockIdx.x);
  const int t_step = BSTAT ? static_cast<int>(gridDim.x) / p.tiles_n : static_cast<int>(gridDim.x);
  const int t_end = BSTAT ? p.tiles_m : num_tiles;
  auto decode = [&](int t) -> TileInfo {
    if constexpr (BSTAT) {
      TileInfo ti;
      const int tm = p.reverse_m ? p.tiles_m - 1 - t : t;
      ti.tile_n = fixed_n; ti.tile_m = tm; ti.split = 0; ti.m0 = tm * kBlockM; ti.n0 = fixed_n * BLOCK_N;
      ti.cta_slot = static_cast<int>(blockIdx.x) / p.tiles_n;
      return ti;
    } else {
      return decode_tile(p, t, BLOCK_N);
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      long long w_empty = 0;
      (void)w_empty;
      auto load_b = [&](uint8_t* sb, uint64_t* bar, int k0, int n0) {
        if constexpr (!B_MN) {
          if (p.b_slab) tma_load_3d(sb, &tmB, bar, 0, n0, k0 >> 6);
          else tma_load_2d(sb, &tmB, bar, k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j) {
            if (p.b_slab) tma_load_3d(sb + j * 8192, &tmB, bar, 0, k0, (n0 >> 6) + j);
            else tma_load_2d(sb + j * 8192, &tmB, bar, n0 + 64 * j, k0);
          }
        }
      };
      if constexpr (BSTAT) {
        if (t_first < t_end) {  // the whole B tile of this CTA's N tile, once
          const int nkb_all = (p.K + kBlockK - 1) / kBlockK;
          mbar_arrive_expect_tx(b_full_bar, nkb_all * Cfg::kBBytes);
          for (int kb = 0; kb < nkb_all; ++kb) load_b(smem + kb * Cfg::kBBytes, b_full_bar, kb * kBlockK, fixed_n * BLOCK_N);
        }
      }
      for (int t = t_first; t < t_end; t += t_step) {
        const TileInfo ti = decode(t);
        const int k_begin = ti.split * p.k_per_split;
        const int k_end = min(p.K, k_begin + p.k_per_split);
        const int nkb = (k_end - k_begin + kBlockK - 1) / kBlockK;
        if constexpr (BSTAT && !A_MN) {
          // The A tile is shared by the tiles_n CTAs of a group and its first touch is an HBM read that queues behind
          // this kernel's own output stream: one CTA of the group pulls it into L2 well ahead of its use.
          const int tp = t + p.a_prefetch * t_step;
          if (p.a_prefetch > 0 && fixed_n == 0 && tp < t_end) {
            for (int kb = 0; kb < nkb; ++kb) {
              if (p.a_slab) tma_prefetch_l2_3d(&tmA, 0, tp * kBlockM, kb);
              else tma_prefetch_l2_2d(&tmA, kb * kBlockK, tp * kBlockM);
            }
          }
        }
        for (int kb = 0; kb < nkb; ++kb) {
          SVB_TRACED_WAIT(w_empty, &empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem_ring + stage * Cfg::kStageBytes;
#ifdef SVB_GEMM_TRACE
          if (p.a_skip > 1 && ((t / t_step) * nkb + kb) % p.a_skip != 0) {  // pretend the tile arrived (e.g. by multicast)
            mbar_arrive(&full_bar[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
#endif
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int k0 = k_begin + kb * kBlockK;
          if constexpr (!A_MN) {
            if (p.a_slab) tma_load_3d(sa, &tmA, &full_bar[stage], 0, ti.m0, k0 >> 6);
            else if (p.a_policy) tma_load_2d_hint(sa, &tmA, &full_bar[stage], k0, ti.m0, p.a_policy);
            else tma_load_2d(sa, &tmA, &full_bar[stage], k0, ti.m0);
          } else {
#pragma unroll
            for (int j = 0; j < kBlockM / 64; ++j) {
              if (p.a_slab) tma_load_3d(sa + j * 8192, &tmA, &full_bar[stage], 0, k0, (ti.m0 >> 6) + j);
              else tma_load_2d(sa + j * 8192, &tmA, &full_bar[stage], ti.m0 + 64 * j, k0);
            }
          }
          if constexpr (!BSTAT) load_b(sa + Cfg::kABytes, &full_bar[stage], k0, ti.n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
#ifdef SVB_GEMM_TRACE
      if (p.trace) p.trace[blockIdx.x * 4 + 0] = w_empty;
#endif
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, A_MN, B_MN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      long long w_full = 0, w_tmem = 0;
      (void)w_full; (void)w_tmem;
      if constexpr (BSTAT) {
        if (t_first < t_end) mbar_wait(b_full_bar, 0);
      }
      for (int t = t_first; t < t_end; t += t_step) {
        const TileInfo ti = decode(t);
        const int k_begin = ti.split * p.k_per_split;
        const int k_end = min(p.K, k_begin + p.k_per_split);
        const int nkb = (k_end - k_begin + kBlockK - 1) / kBlockK;
        SVB_TRACED_WAIT(w_tmem, &tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < nkb; ++kb) {
          SVB_TRACED_WAIT(w_full, &full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem_ring + stage * Cfg::kStageBytes);
          const uint32_t b_base = BSTAT ? smem_u32(smem + kb * Cfg::kBBytes) : a_base + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // K-major: 8-row groups are 1024 B apart (SBO); a k-step is 32 B inside the swizzled 128 B row.
            // MN-major: each 64-wide MN atom holds kBlockK rows of 128 B (LBO = 8192 B between atoms),
            //           8-row K groups are 1024 B apart (SBO); a k-step is 16 rows = 2048 B.
            const uint64_t adesc = A_MN ? make_smem_desc_sw128(a_base + k * 2048, 8192, 1024)
                                        : make_smem_desc_sw128(a_base + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_base + k * 2048, 8192, 1024)
                                        : make_smem_desc_sw128(b_base + k * 32, 16, 1024);
            umma_f16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            if constexpr (ONES) {  // row sums of A: the same A descriptor against the constant ones tile (N = 16)
              constexpr uint32_t idesc1 = make_idesc_bf16(kBlockM, 16, A_MN, false);
              umma_f16(tmem_base + BLOCK_N, adesc, make_smem_desc_sw128(smem_u32(epi_smem), 16, 1024), idesc1,
                       (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);                     // smem slot reusable once these MMAs retire
          if (kb == nkb - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator complete
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (ONES) {
          acc_phase ^= 1;  // single accumulator stage
        } else {
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
#ifdef SVB_GEMM_TRACE
      if (p.trace) { p.trace[blockIdx.x * 4 + 1] = w_full; p.trace[blockIdx.x * 4 + 2] = w_tmem; }
#endif
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 lane quarters x kWarps/4 column groups)
    constexpr int EW = Epi::kWarps;
    static_assert(EW == 8 || EW == 16, "8 or 16 epilogue warps");
    constexpr int kChunksPerWarp = (BLOCK_N / 32) / (EW / 4);
    static_assert(kChunksPerWarp >= 2 && kChunksPerWarp % 2 == 0, "each epilogue warp owns whole 64-column slabs");
    const int ew = warp - 2;  // 0..EW-1
    const int wq = warp % 4;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    const int cgroup = ew / 4;
    const int row_in_tile = wq * 32 + lane;
    const int tid = ew * 32 + lane;
    Epi epi(ep, epi_smem, ew, BLOCK_N);
    const int n_lim = epi_pads_n64<Epi>::value ? ((p.N + 63) & ~63) : p.N;
    uint32_t acc = 0, acc_phase = 0;
    long long w_acc = 0;
    (void)w_acc;
    if (Epi::kColVecs > 0 && t_first < t_end) epi.colvec_fetch(p, decode(t_first), tid);
    for (int t = t_first; t < t_end; t += t_step) {
      const TileInfo ti = decode(t);
      if (Epi::kColVecs > 0) {
        epi.colvec_commit(acc, tid);
        epi_bar_sync(EW * 32);
        if (t + t_step < t_end) epi.colvec_fetch(p, decode(t + t_step), tid);
      }
      SVB_TRACED_WAIT(w_acc, &tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const int row = ti.m0 + row_in_tile;
      epi.begin_tile(p, ti, row, wq, lane);
      const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * BLOCK_N;
      // The warp's chunks are unrolled so that the chunk index `ci` is a compile-time constant inside Epi::chunk
      // (per-chunk state lives in registers without select chains).  Epilogues with kPrefetchAcc read the next
      // chunk's accumulators from TMEM while the current chunk is being processed.
      if constexpr (epi_prefetches_acc<Epi>::value) {
        float v[2][32];
        const int c_begin = cgroup * kChunksPerWarp;
        if (ti.n0 + c_begin * 32 < n_lim) tmem_ld_32x32(t_addr + c_begin * 32, v[0]);
#pragma unroll
        for (int ci = 0; ci < kChunksPerWarp; ++ci) {
          const int c = c_begin + ci;
          const int col0 = ti.n0 + c * 32;
          if (col0 < n_lim) {
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
          if (col0 < n_lim) {
            float v[32];
            if constexpr (!epi_skips_acc_load<Epi>::value) {
              tmem_ld_32x32(t_addr + c * 32, v);
              tmem_ld_wait();
            }
            epi.chunk(p, ti, row, col0, v, wq, lane, ci);
          }
        }
      }
      if constexpr (ONES) {
        if (cgroup == 0) {  // one warp per lane quarter reads the extra column
          const float rs = tmem_ld_32x1(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + BLOCK_N);
          tmem_ld_wait();
          epi.row_sum(p, ti, row, rs);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);  // accumulator stage free for the MMA warp
      epi.end_tile(p, ti, row, wq, lane);
      if constexpr (ONES) {
        acc_phase ^= 1;
      } else {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    epi.finish(wq, lane);  // e.g. drain outstanding bulk stores before the CTA's smem goes away
#ifdef SVB_GEMM_TRACE
    if (p.trace && ew == 0 && lane == 0) p.trace[blockIdx.x * 4 + 3] = w_acc;
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}


}  // namespace svb
