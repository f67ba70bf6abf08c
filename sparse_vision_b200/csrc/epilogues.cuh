// Epilogue functors for gemm_bf16_kernel.  An epilogue thread owns ONE accumulator row (a token, or a feature in the
// weight-gradient GEMMs) and receives its warp's share of the columns 32 at a time (Epi::kWarps = 8 or 16 epilogue warps: warp `ew`
// owns a TMEM lane quarter and column group ew/4 of the tile).  Rows >= M and columns >= N hold zeros from TMA
// out-of-bounds fill, but they must still be masked out of every reduction and store.
//
// bf16 outputs on the hot path leave through shared memory: each warp stages 32-row x 64-column slabs (128-byte rows,
// 128B swizzle, bank-conflict free) and one lane issues a TMA tensor store, so global writes are full 128-byte lines
// instead of 32 scattered 16-byte pieces per instruction (which made the first version of the encoder GEMM 4x
// slower than its MMA time).  All floating-point reductions are written as per-(tile,warp) partials and summed later
// in a fixed order, so a step is bit-reproducible.
#pragma once
#include <type_traits>
#include "gemm_sm100.cuh"

namespace svb {

// 1-bit ReLU masks: word w of row r (32 columns each) lives at [(w / 4) * rows + r] * 4 + w % 4, i.e. the four words
// (128 columns) that one epilogue warp produces for a row are 16 contiguous bytes and the 32 rows of the warp 512
// contiguous bytes -- full-line stores and loads instead of 16-byte pieces at a pitch of words*4 bytes.  w must be a
// multiple of 4 here (a warp's first word).  Buffer size: ceil(words / 4) * rows * 4 words.
__host__ __device__ __forceinline__ size_t mask_index(long long row, int w, long long rows) {
  return (static_cast<size_t>(w >> 2) * static_cast<size_t>(rows) + static_cast<size_t>(row)) * 4 + (w & 3);
}

// Element offset of (row, col) in a slab-major [cols/64][rows][64] matrix (see make_tmap_bf16_slab in gemm_host.cuh).
__host__ __device__ __forceinline__ size_t slab_offset(long long row, int col, long long rows) {
  return (static_cast<size_t>(col >> 6) * static_cast<size_t>(rows) + static_cast<size_t>(row)) * 64 + (col & 63);
}

__device__ __forceinline__ void store_row_f32(float* dst, const float (&v)[32], int nvalid) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i * 4 < nvalid) reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  uint4 q;
  q.x = pack_bf16x2(v[0], v[1]);
  q.y = pack_bf16x2(v[2], v[3]);
  q.z = pack_bf16x2(v[4], v[5]);
  q.w = pack_bf16x2(v[6], v[7]);
  return q;
}
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, const float (&v)[32], int nvalid) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i * 8 < nvalid) reinterpret_cast<uint4*>(dst)[i] = pack8_bf16(v + 8 * i);
}
__device__ __forceinline__ void load_row_bf16(const __nv_bfloat16* src, float (&o)[32], int nvalid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 q = make_uint4(0, 0, 0, 0);
    if (i * 8 < nvalid) q = __ldg(reinterpret_cast<const uint4*>(src) + i);
    o[8 * i + 0] = bf16lo(q.x); o[8 * i + 1] = bf16hi(q.x);
    o[8 * i + 2] = bf16lo(q.y); o[8 * i + 3] = bf16hi(q.y);
    o[8 * i + 4] = bf16lo(q.z); o[8 * i + 5] = bf16hi(q.z);
    o[8 * i + 6] = bf16lo(q.w); o[8 * i + 7] = bf16hi(q.w);
  }
}
__device__ __forceinline__ void lds_row_f32(const float* src, float (&o)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 q = reinterpret_cast<const float4*>(src)[i];
    o[4 * i] = q.x; o[4 * i + 1] = q.y; o[4 * i + 2] = q.z; o[4 * i + 3] = q.w;
  }
}

// Per-warp staging of 32-row x 64-column bf16 slabs that leave through TMA tensor stores.  NBUF = 2 double-buffers
// (a slab is written while the previous one is still being read by the TMA engine); NBUF = 1 waits for the read
// (free when a warp produces one slab per tile).
template <int NBUF>
struct SlabWriterT {
  static constexpr uint32_t kBytesPerWarp = NBUF * 4096;
  __host__ __device__ static constexpr uint32_t bytes(int warps) { return warps * kBytesPerWarp; }
  uint8_t* base;
  uint32_t which;
  bool half_pending;
  __device__ void init(uint8_t* epi_smem, int ew) {
    base = epi_smem + ew * kBytesPerWarp;
    which = 0;
    half_pending = false;
  }
  // columns [half*32, half*32+32) of this lane's row; 16-byte pieces land XOR-swizzled like the tensor map expects
  __device__ void put(int half, int lane, const float (&v)[32]) {
    if (half == 0) {  // the store that last used this buffer must have finished READING it
      if (lane == 0) bulk_wait_read<NBUF - 1>();
      __syncwarp();
    }
    uint8_t* row = base + which * 4096 + lane * 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int j = half * 4 + i;
      *reinterpret_cast<uint4*>(row + ((j ^ (lane & 7)) << 4)) = pack8_bf16(v + 8 * i);
    }
    half_pending = (half == 0);
  }
  // col0 / row0: element coordinates of the slab's first column / row in the output tensor; slab_major: the
  // tensor map is the 3-D slab-major one (gemm_host.cuh), coordinates {0, row, col / 64}
  __device__ void flush(const CUtensorMap* tm, int col0, int row0, int lane, int slab_major = 0) {
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (slab_major) tma_store_3d(tm, base + which * 4096, 0, row0, col0 >> 6);
      else tma_store_2d(tm, base + which * 4096, col0, row0);
      bulk_commit();
    }
    if (NBUF > 1) which ^= 1;
    half_pending = false;
  }
  __device__ void drain(int lane) {
    if (lane == 0) bulk_wait<0>();
    __syncwarp();
  }
};
typedef SlabWriterT<1> SlabWriter1;

// Per-warp staging of ONE 32-row x 32-column bf16 chunk (2 KB, 64-byte rows, 64B swizzle) that leaves through a TMA
// tensor store.  Half the shared memory of a slab writer: with the 128 KB resident weight tile of the B-stationary
// GEMMs that is the difference between a 4-stage and a 5-stage A ring, and ring depth is what those GEMMs are short of
// (0.193 -> 0.171 ms in the probe).  Tensor maps: make_store_tmap_bf16_chunk (row-major, {col, row})
// or make_store_tmap_bf16_slab32 (slab-major, {col % 64, row, col / 64}).
struct ChunkWriter {
  static constexpr uint32_t kBytesPerWarp = 2048;
  __host__ __device__ static constexpr uint32_t bytes(int warps) { return warps * kBytesPerWarp; }
  uint8_t* base;
  __device__ void init(uint8_t* epi_smem, int ew) { base = epi_smem + ew * kBytesPerWarp; }
  // this lane's row of the chunk; the store that last used the tile must have finished READING it
  __device__ __forceinline__ void put(int lane, const float (&v)[32]) {
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
    uint8_t* row = base + lane * 64;
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(row + ((i ^ ((lane >> 1) & 3)) << 4)) = pack8_bf16(v + 8 * i);
  }
  // bf16 element (r, c) of the staged chunk
  __device__ __forceinline__ const uint8_t* at(int r, int c) const {
    return base + r * 64 + ((((c >> 3) ^ ((r >> 1) & 3)) << 4) | ((c & 7) << 1));
  }
  __device__ __forceinline__ void flush(const CUtensorMap* tm, int col0, int row0, int lane, int slab_major) {
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (slab_major) tma_store_3d(tm, base, col0 & 63, row0, col0 >> 6);
      else tma_store_2d(tm, base, col0, row0);
      bulk_commit();
    }
  }
  __device__ void drain(int lane) {
    if (lane == 0) bulk_wait<0>();
    __syncwarp();
  }
};

// The same with NBUF rotating 2 KB tiles per warp: put() only waits for the store issued NBUF chunks ago.  For kernels
// that have the shared memory to spare (the two-CTA B-stationary GEMM keeps only half a weight tile per SM), so that an
// epilogue warp never stalls on the TMA engine's read of its previous chunk inside a tile.
template <int NBUF>
struct ChunkWriterN {
  static constexpr uint32_t kBytesPerWarp = NBUF * 2048;
  __host__ __device__ static constexpr uint32_t bytes(int warps) { return warps * kBytesPerWarp; }
  uint8_t* base0;
  uint8_t* base;
  int which;
  __device__ void init(uint8_t* epi_smem, int ew) { base0 = epi_smem + ew * kBytesPerWarp; base = base0; which = 0; }
  __device__ __forceinline__ void put(int lane, const float (&v)[32]) {
    if (lane == 0) bulk_wait_read<NBUF - 1>();
    __syncwarp();
    uint8_t* row = base + lane * 64;
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(row + ((i ^ ((lane >> 1) & 3)) << 4)) = pack8_bf16(v + 8 * i);
  }
  __device__ __forceinline__ void flush(const CUtensorMap* tm, int col0, int row0, int lane, int slab_major) {
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (slab_major) tma_store_3d(tm, base, col0 & 63, row0, col0 >> 6);
      else tma_store_2d(tm, base, col0, row0);
      bulk_commit();
    }
    which = which + 1 == NBUF ? 0 : which + 1;
    base = base0 + which * 2048;
  }
  __device__ void drain(int lane) {
    if (lane == 0) bulk_wait<0>();
    __syncwarp();
  }
};

// ------------------------------------------------------------------------------------------------ fp32 partials
// Split-K slices of the weight-gradient GEMMs: out[split][row][col] = acc (fp32, direct 16-byte stores; the
// epilogue is a negligible part of these K = T GEMMs, so no staging and a full 4-stage operand ring).
struct EpiPartial {
  struct Params {
    float* out;
    long long ld;
    long long split_stride;
  };
  static constexpr int kWarps = 8;
  static constexpr int kColVecs = 0;
  static constexpr uint32_t kSmemBytes = 0;
  const Params& p;
  __device__ EpiPartial(const Params& p_, uint8_t*, int, int) : p(p_) {}
  __device__ void colvec_fetch(const GemmProblem&, const TileInfo&, int) {}
  __device__ void colvec_commit(uint32_t, int) {}
  __device__ void begin_tile(const GemmProblem&, const TileInfo&, int, int, int) {}
  __device__ __forceinline__ void chunk(const GemmProblem& g, const TileInfo& ti, int row, int col0, float (&v)[32],
                                        int, int, int) {
    if (row >= g.M) return;
    store_row_f32(p.out + ti.split * p.split_stride + static_cast<long long>(row) * p.ld + col0, v,
                  min(32, g.N - col0));
  }
  __device__ void end_tile(const GemmProblem&, const TileInfo&, int, int, int) {}
  __device__ void finish(int, int) {}
};

// Split-K partials plus the row sums of A over this split's K range (kOnesCol, see gemm_sm100.cuh):
// row_out[split][row] = sum_k A[row, k].  In the dW_enc GEMM A = dPre'^T, so these are the per-feature column sums of
// dPre' (-> db_enc and the rank-1 fix-up) at the price of one N = 16 MMA per k-step -- the dE epilogue, which is the
// one that has no cycles to spare, computes none.
struct EpiPartialOnes {
  struct Params {
    float* out;
    long long ld;
    long long split_stride;
    float* row_out;  // [k_splits][M]
  };
  static constexpr int kWarps = 8;
  static constexpr int kColVecs = 0;
  static constexpr bool kOnesCol = true;
  static constexpr uint32_t kSmemBytes = 2048;  // the constant ones tile
  const Params& p;
  __device__ EpiPartialOnes(const Params& p_, uint8_t*, int, int) : p(p_) {}
  __device__ void colvec_fetch(const GemmProblem&, const TileInfo&, int) {}
  __device__ void colvec_commit(uint32_t, int) {}
  __device__ void begin_tile(const GemmProblem&, const TileInfo&, int, int, int) {}
  __device__ __forceinline__ void chunk(const GemmProblem& g, const TileInfo& ti, int row, int col0, float (&v)[32],
                                        int, int, int) {
    if (row >= g.M) return;
    store_row_f32(p.out + ti.split * p.split_stride + static_cast<long long>(row) * p.ld + col0, v,
                  min(32, g.N - col0));
  }
  __device__ void row_sum(const GemmProblem& g, const TileInfo& ti, int row, float v) {
    if (row < g.M) p.row_out[static_cast<size_t>(ti.split) * g.M + row] = v;
  }
  __device__ void end_tile(const GemmProblem&, const TileInfo&, int, int, int) {}
  __device__ void finish(int, int) {}
};

// ------------------------------------------------------------------------------------------------ plain store
// out = [relu](alpha*acc + bias), fp32 (direct) or bf16 (TMA slabs when tm_valid).
struct EpiStore {
  struct Params {
    alignas(64) CUtensorMap tm;  // bf16 output map (box 64 x 32) when tm_valid
    void* out;
    long long ld;
    const float* bias;  // [N] or null
    float alpha;
    int relu;
    int out_bf16;
    int tm_valid;
  };
  static constexpr int kWarps = 8;
  static constexpr int kColVecs = 1;
  static constexpr uint32_t kSmemBytes = SlabWriter1::bytes(kWarps) + 2 * 256 * sizeof(float);
  const Params& p;
  SlabWriter1 slab;
  ColVecStage<1, kWarps * 32> stage;
  float* cv_base;
  const float* cv;
  __device__ EpiStore(const Params& p_, uint8_t* smem, int ew, int)
      : p(p_), cv_base(reinterpret_cast<float*>(smem + SlabWriter1::bytes(kWarps))), cv(cv_base) {
    slab.init(smem, ew);
  }
  __device__ void colvec_fetch(const GemmProblem& g, const TileInfo& ti, int tid) {
    const float* const src[1] = {p.bias};
    stage.fetch(src, ti.n0, g.N, tid);
  }
  __device__ void colvec_commit(uint32_t parity, int tid) {
    float* dst = cv_base + parity * 256;
    stage.commit(dst, tid);
    cv = dst;
  }
  __device__ void begin_tile(const GemmProblem&, const TileInfo&, int, int, int) {}
  __device__ __forceinline__ void chunk(const GemmProblem& g, const TileInfo& ti, int row, int col0, float (&v)[32],
                                        int wq, int lane, int) {
    const int nvalid = min(32, g.N - col0);
    float b[32];
    lds_row_f32(cv + (col0 - ti.n0), b);  // zeros when there is no bias
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] * p.alpha + b[j];
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (p.out_bf16 && p.tm_valid) {
      const int half = (col0 >> 5) & 1;
      slab.put(half, lane, v);
      if (half == 1) slab.flush(&p.tm, col0 - 32, ti.m0 + wq * 32, lane);
      return;
    }
    if (row >= g.M) return;
    const long long off = static_cast<long long>(row) * p.ld + col0;
    if (p.out_bf16) store_row_bf16(reinterpret_cast<__nv_bfloat16*>(p.out) + off, v, nvalid);
    else store_row_f32(reinterpret_cast<float*>(p.out) + off, v, nvalid);
  }
  __device__ void end_tile(const GemmProblem& g, const TileInfo& ti, int, int wq, int lane) {
    if (slab.half_pending) slab.flush(&p.tm, ((g.N - 1) >> 6) << 6, ti.m0 + wq * 32, lane);
  }
  __device__ void finish(int, int lane) { slab.drain(lane); }
};

// ------------------------------------------------------------------------------------------------ encoder
// pre = acc + bias';  e = relu(pre)   (sae_mlp.py:49-51 with the pre-bias folded: bias' = b_enc - W_enc b_dec)
// Fused: bf16 store of e (TMA chunks), optional fp32 stores of e / pre (API forward), per-row mask words (1 bit per
// element: the ReLU mask the backward needs, and what the per-image activity bits of utils.py:2033-2047 are derived
// from by mask_to_activity_kernel) and sum|e| partials (sparse_loss.py:41).
// API = true additionally offers fp32 stores of e / pre (svb_sae_forward); the training step instantiates API = false.
// NBUF > 1: rotating chunk staging (ChunkWriterN) for the two-CTA encoder GEMM.
// WARPS = 16: four epilogue warps per TMEM lane quarter, two chunks each.  The epilogue is ~260 instructions per 32 x 32
// chunk and with two warps per scheduler it issues only half of the cycles (ncu: issue active 50 %, stalls = fixed-latency
// waits and scoreboards), which is what bounds the K = 256 encoder GEMM (4.5 kcycles per tile for 1 kcycle of MMA).
template <bool API, int NBUF = 1, int WARPS = 8>
struct EpiEncT {
  using Writer = typename std::conditional<NBUF == 1, ChunkWriter, ChunkWriterN<NBUF>>::type;
  struct Params {
    alignas(64) CUtensorMap tm_e;  // bf16 e [M,N], 32 x 32 chunks (make_store_tmap_bf16_chunk / _slab32; valid when e_bf16 != null)
    const float* bias;             // [N]
    __nv_bfloat16* e_bf16;         // [M,N] or null
    float* e_f32;                  // [M,N] or null   (API only)
    float* pre_f32;                // [M,N] or null   (API only)
    uint32_t* mask_words;          // group-major (mask_index) or null: bit j of word w <=> e[row, 32w+j] > 0
    float* l1_partial;             // [gridDim.x * kWarps] or null: one running sum per CTA and epilogue warp
    int words;                     // ceil(N/32)
    int e_slab;                    // e_bf16 / tm_e are slab-major (gemm_host.cuh)
  };
  static constexpr int kWarps = WARPS;
  static constexpr int kColVecs = 1;
  static constexpr bool kPrefetchAcc = true;
  static constexpr uint32_t kSmemBytes = Writer::bytes(kWarps) + 2 * 256 * sizeof(float);
  const Params& p;
  Writer slab;
  ColVecStage<1, kWarps * 32> stage;
  float* cv_base;
  const float* cv;
  float sum, total;
  uint32_t words[4];
  int ew, cpw, c_first;  // chunks per warp, this warp's first 32-column chunk inside the tile
  __device__ EpiEncT(const Params& p_, uint8_t* smem, int ew_, int block_n)
      : p(p_), cv_base(reinterpret_cast<float*>(smem + Writer::bytes(kWarps))), cv(cv_base), sum(0.f), total(0.f), ew(ew_),
        cpw((block_n / 32) / (kWarps / 4)), c_first((ew_ / 4) * ((block_n / 32) / (kWarps / 4))) {
    slab.init(smem, ew_);
  }
  __device__ void colvec_fetch(const GemmProblem& g, const TileInfo& ti, int tid) {
    const float* const src[1] = {p.bias};
    stage.fetch(src, ti.n0, g.N, tid);
  }
  __device__ void colvec_commit(uint32_t parity, int tid) {
    float* dst = cv_base + parity * 256;
#pragma unroll
    for (int i = 0; i < decltype(stage)::kPer; ++i) stage.r[i] = 0.f - stage.r[i];  // stage -bias' (+0 stays +0)
    stage.commit(dst, tid);
    cv = dst;
  }
  __device__ void begin_tile(const GemmProblem&, const TileInfo&, int, int, int) {
    sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) words[i] = 0;
  }
  // ci: index of the chunk among this warp's chunks (a compile-time constant after the caller's unrolling)
  __device__ __forceinline__ void chunk(const GemmProblem& g, const TileInfo& ti, int row, int col0, float (&v)[32],
                                        int wq, int lane, int ci) {
    const int nvalid = min(32, g.N - col0);
    const bool row_ok = row < g.M;
    float nb[32];
    lds_row_f32(cv + (col0 - ti.n0), nb);  // -bias'
    const long long off = static_cast<long long>(row) * g.N + col0;
    if (API && p.pre_f32) {                // materialise pre = acc + bias'
      float pre[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) pre[j] = v[j] - nb[j];
      if (row_ok) store_row_f32(p.pre_f32 + off, pre, nvalid);
    }
    // t = (-bias') - acc = -pre: its sign bit is set exactly when pre > 0 (pre == +-0 gives +0), e = max(-t, 0).
    // Four independent chains (8 columns each) for the mask bits and the partial sums.
    uint32_t wq4[4] = {0, 0, 0, 0};
    float sq4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 7; j >= 0; --j) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = q * 8 + j;
        const float t = nb[i] - v[i];
        wq4[q] = __funnelshift_l(__float_as_uint(t), wq4[q], 1);
        v[i] = fmaxf(-t, 0.f);
        sq4[q] += v[i];
      }
    }
    uint32_t word = (wq4[0] | (wq4[1] << 8)) | ((wq4[2] << 16) | (wq4[3] << 24));
    sum += (sq4[0] + sq4[1]) + (sq4[2] + sq4[3]);
    if (nvalid < 32) word &= (1u << nvalid) - 1u;
    if (!row_ok) word = 0;
    words[ci] = word;
    if (p.e_bf16) {
      slab.put(lane, v);
      slab.flush(&p.tm_e, col0, ti.m0 + wq * 32, lane, p.e_slab);
    }
    if (API && p.e_f32 && row_ok) store_row_f32(p.e_f32 + off, v, nvalid);
  }
  __device__ void end_tile(const GemmProblem& g, const TileInfo& ti, int row, int wq, int lane) {
    const int w0 = (ti.n0 >> 5) + c_first;            // first word index of this warp
    const int nw = max(0, min(cpw, p.words - w0));
    if (p.mask_words && row < g.M && nw > 0) {
      // group-major layout [ceil(words/4)][M][4]: the 32 rows of a warp are 512 contiguous bytes (see mask_index)
      uint32_t* dst = p.mask_words + mask_index(row, w0, g.M);
      if (nw == 4) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(words[0], words[1], words[2], words[3]);
      } else if (nw == 2 && cpw == 2) {   // 16 warps: a warp's two words are 8 contiguous, 8-byte aligned bytes
        *reinterpret_cast<uint2*>(dst) = make_uint2(words[0], words[1]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < nw) dst[i] = words[i];
      }
    }
    if (row < g.M) total += sum;  // rows >= M hold relu(bias'), not data
  }
  __device__ void finish(int, int lane) {
    slab.drain(lane);
    if (p.l1_partial) {
      const float s = warp_sum(total);
      if (lane == 0) p.l1_partial[static_cast<size_t>(blockIdx.x) * kWarps + ew] = s;
    }
  }
};
typedef EpiEncT<false> EpiEnc;     // training step
typedef EpiEncT<false, 4> EpiEnc4; // training step, two-CTA GEMM: 4 rotating chunk tiles per warp
typedef EpiEncT<false, 1, 16> EpiEnc16;  // training step, B-stationary GEMM with 16 epilogue warps
typedef EpiEncT<true> EpiEncApi;   // svb_sae_forward (optional fp32 e / pre outputs)

// ------------------------------------------------------------------------------------------------ decoder
// d = acc + b_dec;  diff = d - x   (sae_mlp.py:52, sparse_loss.py:35).  Fused: stores of d / diff, sum diff^2.
// HBM-bound (it streams E once), so 8 epilogue warps are enough.
struct EpiDec {
  struct Params {
    alignas(64) CUtensorMap tm_d;     // bf16 d [M,N]   (valid when d_bf16 != null)
    alignas(64) CUtensorMap tm_diff;  // bf16 diff [M,N] (valid when diff_bf16 != null)
    const float* bias;                // [N] decoder bias
    const __nv_bfloat16* x;           // [M,N] targets (SAE input) or null
    __nv_bfloat16* d_bf16;            // [M,N] or null
    float* d_f32;                     // [M,N] or null
    __nv_bfloat16* diff_bf16;         // [M,N] or null
    float* sq_partial;                // [gridDim.x * kWarps] or null: one running sum per CTA and epilogue warp
    int out_slab;                     // d_bf16 / diff_bf16 and their maps are slab-major
    int x_slab;                       // x is slab-major
  };
  static constexpr int kWarps = 8;
  static constexpr int kColVecs = 1;
  static constexpr uint32_t kSmemBytes = 2 * SlabWriter1::bytes(kWarps) + 2 * 256 * sizeof(float);
  const Params& p;
  SlabWriter1 slab_d, slab_f;
  ColVecStage<1, kWarps * 32> stage;
  float* cv_base;
  const float* cv;
  float sq;  // running over all tiles of this CTA (rows >= M and columns >= N never enter it)
  int ew;
  __device__ EpiDec(const Params& p_, uint8_t* smem, int ew_, int)
      : p(p_), cv_base(reinterpret_cast<float*>(smem + 2 * SlabWriter1::bytes(kWarps))), cv(cv_base), sq(0.f),
        ew(ew_) {
    slab_d.init(smem, ew_);
    slab_f.init(smem + SlabWriter1::bytes(kWarps), ew_);
  }
  __device__ void colvec_fetch(const GemmProblem& g, const TileInfo& ti, int tid) {
    const float* const src[1] = {p.bias};
    stage.fetch(src, ti.n0, g.N, tid);
  }
  __device__ void colvec_commit(uint32_t parity, int tid) {
    float* dst = cv_base + parity * 256;
    stage.commit(dst, tid);
    cv = dst;
  }
  __device__ void begin_tile(const GemmProblem&, const TileInfo&, int, int, int) {}
  __device__ __forceinline__ void chunk(const GemmProblem& g, const TileInfo& ti, int row, int col0, float (&v)[32],
                                        int wq, int lane, int) {
    const int nvalid = min(32, g.N - col0);
    const bool row_ok = row < g.M;
    const int half = ((col0 - ti.n0) >> 5) & 1;
    float b[32];
    const long long off = static_cast<long long>(row) * g.N + col0;
    float xv[32];
    if (p.x)  // issued early: an L2 round trip
      load_row_bf16(p.x + (row_ok ? (p.x_slab ? static_cast<long long>(slab_offset(row, col0, g.M)) : off) : 0), xv,
                    row_ok ? nvalid : 0);
    lds_row_f32(cv + (col0 - ti.n0), b);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += b[j];
    if (p.d_bf16) {
      slab_d.put(half, lane, v);
      if (half == 1) slab_d.flush(&p.tm_d, col0 - 32, ti.m0 + wq * 32, lane, p.out_slab);
    }
    if (p.d_f32 && row_ok) store_row_f32(p.d_f32 + off, v, nvalid);
    if (p.x) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] -= xv[j];
        if (j < nvalid && row_ok) sq += v[j] * v[j];
      }
      if (p.diff_bf16) {
        slab_f.put(half, lane, v);
        if (half == 1) slab_f.flush(&p.tm_diff, col0 - 32, ti.m0 + wq * 32, lane, p.out_slab);
      }
    }
  }
  __device__ void end_tile(const GemmProblem& g, const TileInfo& ti, int, int wq, int lane) {
    const int last = ((g.N - 1) >> 6) << 6;
    if (slab_d.half_pending) slab_d.flush(&p.tm_d, last, ti.m0 + wq * 32, lane, p.out_slab);
    if (slab_f.half_pending) slab_f.flush(&p.tm_diff, last, ti.m0 + wq * 32, lane, p.out_slab);
  }
  __device__ void finish(int, int lane) {
    slab_d.drain(lane);
    if (p.sq_partial) {
      const float s = warp_sum(sq);
      if (lane == 0) p.sq_partial[static_cast<size_t>(blockIdx.x) * kWarps + ew] = s;
    }
  }
};

// ------------------------------------------------------------------------------------------------ decoder, fused
// Training-step decoder epilogue for NCHW activations (slab-major X / DIFF, C % 64 == 0, >= 32 tokens per image):
//   d = acc + b_dec;  diff = d - x;  sum diff^2                                  (sae_mlp.py:52, sparse_loss.py:35)
// and, from two small staged tiles per 32-column chunk, everything the separate post-decoder pass did:
//   * d goes straight back to the caller's NCHW bf16 tensor (model_pipeline.py:425,432; utils.py:2478): the warp
//     stages a channel-major [32 channels][32 tokens] copy and one TMA tensor store writes it (positions past the end
//     of the image are clipped by the tensor map; TMA stores reject negative coordinates, so when the warp's 32
//     tokens straddle an image boundary the few positions of the second image are copied with 16-byte stores);
//   * per 32-token group and channel: sum d, sum d^2 per image (variance_explained, utils.py:2012-2030) read back
//     from the channel-major copy, and sum diff^2 (compute_rmse_nrmse, sparse_loss.py:4-21) read back from the diff
//     tile.  dec_stats_image_kernel folds the groups of an image together.  No token-major d is written at all.
// Staging is per chunk (2 x 2 KB per warp) on purpose: 34 KB of epilogue smem leave room for FOUR 48 KB operand
// stages, and this GEMM streams E from HBM -- with three stages it ran at 4.1 TB/s, with four at 4.9 TB/s.
struct EpiDecNchwParams {
  alignas(64) CUtensorMap tm_diff;  // slab-major bf16 diff [M, N], box 32 cols x 32 rows x 1, 64B swizzle
  alignas(64) CUtensorMap tm_out;   // NCHW bf16 output seen as {HW, C, B}, box {32, 32, 1}, no swizzle (out != null)
  const float* bias;                // [N] decoder bias
  const __nv_bfloat16* x;           // slab-major [M, N] targets (the SAE input)
  float* sq_partial;                // [gridDim.x * kWarps]: one running sum per CTA and epilogue warp
  float* part;                      // [(tiles_m * 4 groups) * 2 slots][3][N], see dec_stats_image_kernel
  void* out;                        // out_kind 1: the caller's NCHW bf16 tensor (second-image pieces of straddling warps)
  int hw;                           // tokens per image (>= 32)
  int out_kind;                     // 1: bf16 through tm_out (HW % 8 == 0, 16-byte aligned base);
                                    // 4: tm_out is a channel-major [C][T] workspace (make_store_tmap_bf16_cmajor)
                                    //    that a copy kernel turns into the caller's tensor (any HW, bf16 or fp32)
                                    // 2: token-major bf16 [M, N] through tm_out (make_store_tmap_bf16_chunk): the
                                    //    caller's channels_last tensor, no layout change at all (tok = 1 only)
  int x_slab;                       // x is slab-major (1) or the caller's row-major token matrix [M, N] (0)
  int tok;                          // 1: token-major in / out: d is staged token-major, the per-channel sums are
                                    //    read back column-wise like those of diff (no channel-major copy exists)
};
// WARPS = 8: epilogue of gemm_bf16_kernel; WARPS = 16: decoder epilogue of the fused forward kernel (fused_fwd_sm100.cuh),
// which also places the bias vector itself (use_colvec_at) because 16 x 4 KB of staging fill the aliased E tile.
template <int WARPS>
struct EpiDecNchwT {
  using Params = EpiDecNchwParams;
  static constexpr int kWarps = WARPS;
  static constexpr int kColVecs = 1;
  static constexpr bool kPrefetchAcc = true;
  static constexpr bool kPadN64 = true;   // N % 64 != 0: the padding columns of diff's last slab are written (zeros)
  static constexpr uint32_t kSmemBytes = kWarps * 4096 + 2 * 256 * sizeof(float);
  const Params& p;
  uint8_t* tbuf;  // channel-major [32 channels][32 tokens] bf16 copy of d (2 KB)
  uint8_t* fbuf;  // token-major [32 tokens][32 channels] bf16 diff, 64B-swizzled (2 KB)
  ColVecStage<1, kWarps * 32> stage;
  float* cv_base;
  const float* cv;
  float sq;
  int ew;
  __device__ EpiDecNchwT(const Params& p_, uint8_t* smem, int ew_, int)
      : p(p_), tbuf(smem + ew_ * 4096), fbuf(smem + ew_ * 4096 + 2048),
        cv_base(reinterpret_cast<float*>(smem + kWarps * 4096)), cv(cv_base), sq(0.f), ew(ew_) {}
  __device__ void use_colvec_at(float* dst) { cv_base = dst; cv = dst; }   // colvec_commit(0, tid) then writes there
  __device__ void colvec_fetch(const GemmProblem& g, const TileInfo& ti, int tid) {
    const float* const src[1] = {p.bias};
    stage.fetch(src, ti.n0, g.N, tid);
  }
  __device__ void colvec_commit(uint32_t parity, int tid) {
    float* dst = cv_base + parity * 256;
    stage.commit(dst, tid);
    cv = dst;
  }
  __device__ void begin_tile(const GemmProblem&, const TileInfo&, int, int, int) {}
  __device__ __forceinline__ void chunk(const GemmProblem& g, const TileInfo& ti, int row, int col0, float (&v)[32],
                                        int wq, int lane, int) {
    const bool row_ok = row < g.M;
    float xv[32], b[32];
    if (p.x_slab)  // issued early; the padding columns of the last slab exist (zeros)
      load_row_bf16(p.x + (row_ok ? slab_offset(row, col0, g.M) : 0), xv, row_ok ? 32 : 0);
    else           // row-major [M, N]: columns >= N (padding chunks of the last diff slab) are not there
      load_row_bf16(p.x + (row_ok ? static_cast<size_t>(row) * g.N + min(col0, g.N - 8) : 0), xv,
                    row_ok ? max(0, min(32, g.N - col0)) : 0);
    lds_row_f32(cv + (col0 - ti.n0), b);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += b[j];
    if (!row_ok) {  // only in a ragged last tile: keep padding rows out of the stores and the statistics
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    uint32_t dpk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) dpk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] -= xv[j];
      sq += v[j] * v[j];
    }
    // the stores of the previous chunk must have finished READING the two staging tiles
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
    if (p.tok) {  // token-major [32 tokens][32 channels], 64B-swizzled like the diff tile
      uint8_t* drow = tbuf + lane * 64;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint4*>(drow + ((i ^ ((lane >> 1) & 3)) << 4)) = make_uint4(dpk[4 * i], dpk[4 * i + 1], dpk[4 * i + 2], dpk[4 * i + 3]);
    } else {
      uint16_t* tb = reinterpret_cast<uint16_t*>(tbuf) + lane;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        tb[(2 * j) * 32] = static_cast<uint16_t>(dpk[j] & 0xFFFFu);
        tb[(2 * j + 1) * 32] = static_cast<uint16_t>(dpk[j] >> 16);
      }
    }
    uint8_t* frow = fbuf + lane * 64;
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(frow + ((i ^ ((lane >> 1) & 3)) << 4)) = pack8_bf16(v + 8 * i);
    __syncwarp();  // both staged tiles are complete in shared memory

    const int row0 = ti.m0 + wq * 32;
    const int nrows = max(0, min(32, g.M - row0));
    const int b0 = row0 / p.hw;
    const int n0 = min((b0 + 1) * p.hw - row0, nrows), n1 = nrows - n0;  // tokens of image b0 / of image b0 + 1
    const size_t grp = static_cast<size_t>(ti.tile_m) * 4 + wq;
    float* part0 = p.part + (grp * 2 + 0) * 3 * g.N;
    float* part1 = p.part + (grp * 2 + 1) * 3 * g.N;
    const int col = col0 + lane;  // this lane's channel
    const bool col_ok = col < g.N;  // false only in the padding of the last slab (N % 64 != 0)
    if (p.tok) {  // (1) sum d, sum d^2 of channel col0 + lane: its column of the token-major tile
      float a0 = 0.f, q0 = 0.f, a1 = 0.f, q1 = 0.f;
      const uint8_t* cp = tbuf + ((lane & 7) << 1);
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const uint16_t h = *reinterpret_cast<const uint16_t*>(cp + r * 64 + (((lane >> 3) ^ ((r >> 1) & 3)) << 4));
        const float d = __uint_as_float(static_cast<uint32_t>(h) << 16);
        if (r < n0) { a0 += d; q0 += d * d; }
        else { a1 += d; q1 += d * d; }          // rows >= nrows hold zeros
      }
      if (col_ok) {
        part0[col] = a0;
        part0[g.N + col] = q0;
        if (n1 > 0) { part1[col] = a1; part1[g.N + col] = q1; }
      }
    } else {  // (1) sum d, sum d^2 of channel col0 + lane from its row of the channel-major copy
      const uint4* rp = reinterpret_cast<const uint4*>(tbuf + lane * 64);
      uint32_t w[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 t = rp[q];
        w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
      }
      float a0 = 0.f, q0 = 0.f, a1 = 0.f, q1 = 0.f;
      if (n0 == 32) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float lo = bf16lo(w[i]), hi = bf16hi(w[i]);
          a0 += lo; q0 += lo * lo;
          a0 += hi; q0 += hi * hi;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float d = (i & 1) ? bf16hi(w[i >> 1]) : bf16lo(w[i >> 1]);
          if (i < n0) { a0 += d; q0 += d * d; }
          else if (i < nrows) { a1 += d; q1 += d * d; }
        }
      }
      if (col_ok) {
        part0[col] = a0;
        part0[g.N + col] = q0;
        if (n1 > 0) { part1[col] = a1; part1[g.N + col] = q1; }
      }
    }
    {  // (2) sum diff^2 of the same channel from the (64B-swizzled) diff tile; rows >= M hold zeros
      float s = 0.f;
      const uint8_t* cp = fbuf + ((lane & 7) << 1);
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const uint16_t h = *reinterpret_cast<const uint16_t*>(cp + r * 64 + (((lane >> 3) ^ ((r >> 1) & 3)) << 4));
        const float f = __uint_as_float(static_cast<uint32_t>(h) << 16);
        s += f * f;
      }
      if (col_ok) part0[2 * g.N + col] = s;
    }
    // (3) d back to NCHW.  out_kind 1: one TMA store for the positions of image b0 (clipped at the image end); the
    // positions of a second image in a straddling warp (n0, n1 multiples of 8 there) are copied with 16-byte stores.
    // out_kind 4: one TMA store into the channel-major workspace.
    if (p.out_kind == 1) {
      if (n1 > 0 && col_ok) {
        const uint4* src = reinterpret_cast<const uint4*>(tbuf + lane * 64 + n0 * 2);
        uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (static_cast<size_t>(b0 + 1) * g.N + col) * p.hw);
        for (int q = 0; q < (n1 >> 3); ++q) dst[q] = src[q];
      }
    }
    // (4) asynchronous stores of both tiles
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (p.out_kind == 1 && col0 < g.N) tma_store_3d(&p.tm_out, tbuf, row0 - b0 * p.hw, col0, b0);   // clipped at C
      if (p.out_kind == 4 && col0 < g.N) tma_store_2d(&p.tm_out, tbuf, row0, col0);                    // clipped at T, C
      if (p.out_kind == 2 && col0 < g.N) tma_store_2d(&p.tm_out, tbuf, col0, row0);                    // token-major, clipped
      tma_store_3d(&p.tm_diff, fbuf, col0 & 63, row0, col0 >> 6);
      bulk_commit();
    }
  }
  __device__ void end_tile(const GemmProblem&, const TileInfo&, int, int, int) {}
  __device__ void finish(int, int lane) {
    if (lane == 0) bulk_wait<0>();
    __syncwarp();
    if (p.sq_partial) {
      const float s = warp_sum(sq);
      if (lane == 0) p.sq_partial[static_cast<size_t>(blockIdx.x) * kWarps + ew] = s;
    }
  }
};
typedef EpiDecNchwT<8> EpiDecNchw;

// ------------------------------------------------------------------------------------------------ dE -> dPre
// acc = diff * W_dec  (unscaled dE);  dPre' = 1[e>0] * (acc + l1c)  with l1c = lambda*C/(2F), i.e. the whole
// backward is carried in units of T*C/2 and rescaled once in the gradient reduction (model_pipeline.py:385 autograd
// of sparse_loss.py:35,41 through sae_mlp.py:51).  The ReLU mask comes from the encoder's 1-bit activity words
// (8 B per row and warp instead of re-reading 128 B of e).  Fused: bf16 store of dPre' (TMA slabs), per-feature
// column sums (-> db_enc).
// Column sums over tokens (-> db_enc) are read back from the staged 32 x 32 bf16 chunk in shared memory: the two
// half-warps take the even / odd rows, lane l % 16 owns columns 2l, 2l+1, adds its 16 rows 8 at a time as packed
// bf16 pairs (HADD2) and the two groups and the two halves in fp32 -- about one instruction per element instead of four
// for a 31-step shuffle transpose.
// CS = 0: one partial row per (M tile, lane quarter).  CS = 1 (B-stationary launches: a CTA keeps ONE N tile): the
// sums are carried in registers over all M tiles of the CTA and written once.  CS = 2: no column sums.
template <int CS>
struct EpiDPreT {
  struct Params {
    alignas(64) CUtensorMap tm_dpre;   // bf16 dPre' [M,N], 32 x 32 chunks (make_store_tmap_bf16_chunk / _slab32)
    const uint32_t* mask_words;        // group-major 1-bit ReLU masks of the encoder (mask_index)
    float* colsum_partial;             // CS = 1: [slots * 4, N] with slot = TileInfo::cta_slot;
                                       // CS = 0: [tiles_m * 4 lane quarters, N], one row per 32 tokens
    float l1c;
    int words;
    int out_slab;                      // dPre' and tm_dpre are slab-major
  };
  static constexpr int kWarps = 8;
  static constexpr int kColVecs = 0;
  static constexpr uint32_t kSmemBytes = ChunkWriter::bytes(kWarps);
  const Params& p;
  ChunkWriter slab;
  uint32_t words[4];
  float2 csacc[4];  // CS = 1: running sums of this lane's two columns, per chunk of the warp
  int ew, cpw, c_first, n0_last, slot_last, N_last;
  __device__ EpiDPreT(const Params& p_, uint8_t* smem, int ew_, int block_n_)
      : p(p_), ew(ew_), cpw((block_n_ / 32) / (kWarps / 4)), c_first((ew_ / 4) * ((block_n_ / 32) / (kWarps / 4))),
        n0_last(-1), slot_last(0), N_last(0) {
    slab.init(smem, ew_);
#pragma unroll
    for (int i = 0; i < 4; ++i) csacc[i] = make_float2(0.f, 0.f);
  }
  __device__ void colvec_fetch(const GemmProblem&, const TileInfo&, int) {}
  __device__ void colvec_commit(uint32_t, int) {}
  __device__ void begin_tile(const GemmProblem& g, const TileInfo& ti, int row, int, int) {
#pragma unroll
    for (int i = 0; i < 4; ++i) words[i] = 0;
    const int w0 = (ti.n0 >> 5) + c_first;
    const int nw = max(0, min(cpw, p.words - w0));
    if (row < g.M && nw > 0) {
      const uint32_t* src = p.mask_words + mask_index(row, w0, g.M);
      if (nw == 4) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(src));
        words[0] = a.x; words[1] = a.y; words[2] = a.z; words[3] = a.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < nw) words[i] = __ldg(src + i);
      }
    }
  }
  // Sum of the staged chunk's 32 rows for columns 2*(lane%16), +1 (every lane of a pair l, l+16 gets the total).
  __device__ __forceinline__ float2 chunk_colsum(int lane) const {
    const int cp = lane & 15, par = lane >> 4;  // column pair, row parity of this half-warp
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int g8 = 0; g8 < 2; ++g8) {
      __nv_bfloat162 acc;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(slab.at(2 * (g8 * 8 + i) + par, 2 * cp));
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w);
        acc = i == 0 ? h : __hadd2(acc, h);
      }
      const uint32_t aw = *reinterpret_cast<const uint32_t*>(&acc);
      s.x += bf16lo(aw);
      s.y += bf16hi(aw);
    }
    s.x += __shfl_xor_sync(0xffffffffu, s.x, 16);
    s.y += __shfl_xor_sync(0xffffffffu, s.y, 16);
    return s;
  }
  __device__ __forceinline__ void chunk(const GemmProblem& g, const TileInfo& ti, int, int col0, float (&v)[32], int wq,
                                        int lane, int ci) {
    const uint32_t word = words[ci];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (word & (1u << j)) ? v[j] + p.l1c : 0.f;
    slab.put(lane, v);
    if (CS != 2) {
      __syncwarp();  // every lane's row of the chunk is in shared memory
      const float2 cs = chunk_colsum(lane);
      if (CS == 1) {
        csacc[ci].x += cs.x;
        csacc[ci].y += cs.y;
      } else {
        const int col = col0 + 2 * (lane & 15);  // N % 8 == 0: a column pair is inside or outside together
        if (lane < 16 && col < g.N)
          *reinterpret_cast<float2*>(p.colsum_partial + (static_cast<size_t>(ti.tile_m) * 4 + wq) * g.N + col) = cs;
      }
    }
    slab.flush(&p.tm_dpre, col0, ti.m0 + wq * 32, lane, p.out_slab);
  }
  __device__ void end_tile(const GemmProblem& g, const TileInfo& ti, int, int, int) {
    n0_last = ti.n0; slot_last = ti.cta_slot; N_last = g.N;
  }
  __device__ void finish(int wq, int lane) {
    slab.drain(lane);
    if (CS == 1 && n0_last >= 0 && lane < 16) {
      const size_t rowp = static_cast<size_t>(slot_last) * 4 + wq;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int col = n0_last + (c_first + ci) * 32 + 2 * lane;
        if (ci < cpw && col < N_last) *reinterpret_cast<float2*>(p.colsum_partial + rowp * N_last + col) = csacc[ci];
      }
    }
  }
};
typedef EpiDPreT<0> EpiDPre;
typedef EpiDPreT<1> EpiDPreCta;
typedef EpiDPreT<2> EpiDPreNoSum;

}  // namespace svb
