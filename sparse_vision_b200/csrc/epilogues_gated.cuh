// Epilogues of the Gated-SAE GEMMs (models/gated_sae.py:28-56).  See epilogues.cuh for the conventions.
#pragma once
#include "epilogues.cuh"

namespace svb {

// ------------------------------------------------------------------------------------------------ gated encoder
// One GEMM raw = X W_gate^T feeds both sub-layers (gated_sae.py:37-48 with W_mag = exp(r_mag) * W_gate shared):
//   raw' = acc - dot      (dot[f] = W_gate[f,:] . b_dec, the folded pre-bias)
//   pi   = raw' + b_gate;            relu_pi = relu(pi)
//   mag  = relu(exp(r) * raw' + b_mag)
//   e    = heaviside(pi, 0.5) * mag
// Fused: bf16 stores of e and relu_pi (two swizzled smem slabs per warp + TMA tensor stores when the maps are valid,
// else direct / fp32 stores for the API forward), 1-bit masks of e > 0 and relu_pi > 0 for the backward, activity
// bits of e (utils.py:2033-2047), sum|relu_pi| (sparse_loss.py:71) kept per CTA and warp.
struct EpiGatedEnc {
  struct Params {
    alignas(64) CUtensorMap tm_e;   // bf16 e       [M,N] (valid when tma != 0)
    alignas(64) CUtensorMap tm_rp;  // bf16 relu_pi [M,N]
    const float* dot;     // [N]
    const float* b_gate;  // [N]
    const float* b_mag;   // [N]
    const float* exp_r;   // [N]
    __nv_bfloat16* e_bf16;
    float* e_f32;
    __nv_bfloat16* rp_bf16;
    float* rp_f32;
    uint32_t* mask_e;     // group-major (mask_index) or null: bit j of word w <=> e[row, 32w+j] > 0; the per-image
                          // activity bits (utils.py:2033-2047) are derived from it by mask_to_activity_kernel
    uint32_t* mask_rp;    // group-major or null: relu_pi > 0
    float* l1_partial;    // [gridDim.x * kWarps] or null: one running sum per CTA and epilogue warp
    int words;
    int tma;              // 1: e_bf16 / rp_bf16 leave through the TMA maps; slab_major: those maps are slab-major
    int slab_major;
  };
  static constexpr int kWarps = 8;
  static constexpr int kColVecs = 4;
  static constexpr uint32_t kSmemBytes = 2 * SlabWriter1::bytes(kWarps) + 2 * 4 * 256 * sizeof(float);
  const Params& p;
  SlabWriter1 slab_e, slab_r;
  ColVecStage<4, kWarps * 32> stage;
  float* cv_base;
  const float* cv;
  float sum, total;
  uint32_t we[4], wr[4];
  int ew, cpw, c_first;
  __device__ EpiGatedEnc(const Params& p_, uint8_t* smem, int ew_, int block_n)
      : p(p_), cv_base(reinterpret_cast<float*>(smem + 2 * SlabWriter1::bytes(kWarps))), cv(cv_base), sum(0.f),
        total(0.f), ew(ew_), cpw((block_n / 32) / (kWarps / 4)), c_first((ew_ / 4) * ((block_n / 32) / (kWarps / 4))) {
    slab_e.init(smem, ew_);
    slab_r.init(smem + SlabWriter1::bytes(kWarps), ew_);
  }
  __device__ void colvec_fetch(const GemmProblem& g, const TileInfo& ti, int tid) {
    const float* const src[4] = {p.dot, p.b_gate, p.b_mag, p.exp_r};
    stage.fetch(src, ti.n0, g.N, tid);
  }
  // staged per column: c1 = b_gate - dot, c2 = b_mag - exp(r)*dot, exp(r)   (pi = acc + c1, mag_pre = exp(r)*acc + c2);
  // thread tid holds the four source vectors of column tid (ColVecStage layout with 256 threads)
  __device__ void colvec_commit(uint32_t parity, int tid) {
    static_assert(decltype(stage)::kPer == 4, "one column per epilogue thread");
    float* dst = cv_base + parity * 4 * 256;
    const float dot = stage.r[0], bg = stage.r[1], bm = stage.r[2], er = stage.r[3];
    dst[tid] = bg - dot;
    dst[256 + tid] = bm - er * dot;
    dst[512 + tid] = er;
    cv = dst;
  }
  __device__ void begin_tile(const GemmProblem&, const TileInfo&, int, int, int) {
    sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { we[i] = 0; wr[i] = 0; }
  }
  __device__ __forceinline__ void chunk(const GemmProblem& g, const TileInfo& ti, int row, int col0, float (&v)[32],
                                        int wq, int lane, int ci) {
    const int nvalid = min(32, g.N - col0);
    const bool row_ok = row < g.M;
    float rp[32], c1[32], c2[32], er[32];
    const float* cvt = cv + (col0 - ti.n0);
    lds_row_f32(cvt, c1);
    lds_row_f32(cvt + 256, c2);
    lds_row_f32(cvt + 512, er);
    // Columns >= N see acc = 0 and zero column vectors, so they produce e = relu_pi = 0 by themselves.
    // gate = heaviside(pi, 0.5) as one saturating FMA: sat(pi * 2^127 + 0.5) is 1 / 0.5 / 0 for pi > 0 / == 0 / < 0
    // (exact down to |pi| ~ 3e-39).  Mask bits come from the sign of 0 - e and 0 - relu_pi, four chains of 8 columns.
    uint32_t we4[4] = {0, 0, 0, 0}, wr4[4] = {0, 0, 0, 0};
    float sq4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 7; j >= 0; --j) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = q * 8 + j;
        const float pi = v[i] + c1[i];
        const float mag = fmaxf(fmaf(er[i], v[i], c2[i]), 0.f);
        const float gate = __saturatef(fmaf(pi, 1.7014118e38f, 0.5f));
        const float e = gate * mag;
        const float r = fmaxf(pi, 0.f);
        we4[q] = __funnelshift_l(__float_as_uint(0.f - e), we4[q], 1);
        wr4[q] = __funnelshift_l(__float_as_uint(0.f - r), wr4[q], 1);
        sq4[q] += r;
        v[i] = e;
        rp[i] = r;
      }
    }
    uint32_t word_e = (we4[0] | (we4[1] << 8)) | ((we4[2] << 16) | (we4[3] << 24));
    uint32_t word_r = (wr4[0] | (wr4[1] << 8)) | ((wr4[2] << 16) | (wr4[3] << 24));
    sum += (sq4[0] + sq4[1]) + (sq4[2] + sq4[3]);
    if (nvalid < 32) { word_e &= (1u << nvalid) - 1u; word_r &= (1u << nvalid) - 1u; }
    if (!row_ok) { word_e = 0; word_r = 0; }
    we[ci] = word_e;
    wr[ci] = word_r;
    const long long off = static_cast<long long>(row) * g.N + col0;
    if (p.tma) {
      const int half = ci & 1;
      slab_e.put(half, lane, v);
      slab_r.put(half, lane, rp);
      if (half == 1) {
        slab_e.flush(&p.tm_e, col0 - 32, ti.m0 + wq * 32, lane, p.slab_major);
        slab_r.flush(&p.tm_rp, col0 - 32, ti.m0 + wq * 32, lane, p.slab_major);
      }
    } else if (row_ok) {
      if (p.e_bf16) store_row_bf16(p.e_bf16 + off, v, nvalid);
      if (p.rp_bf16) store_row_bf16(p.rp_bf16 + off, rp, nvalid);
    }
    if (row_ok) {
      if (p.e_f32) store_row_f32(p.e_f32 + off, v, nvalid);
      if (p.rp_f32) store_row_f32(p.rp_f32 + off, rp, nvalid);
    }
  }
  __device__ void end_tile(const GemmProblem& g, const TileInfo& ti, int row, int wq, int lane) {
    if (p.tma) {
      const int last = ((g.N - 1) >> 6) << 6;
      if (slab_e.half_pending) slab_e.flush(&p.tm_e, last, ti.m0 + wq * 32, lane, p.slab_major);
      if (slab_r.half_pending) slab_r.flush(&p.tm_rp, last, ti.m0 + wq * 32, lane, p.slab_major);
    }
    const int w0 = (ti.n0 >> 5) + c_first;
    const int nw = max(0, min(cpw, p.words - w0));
    if (row < g.M && nw > 0) {
      const size_t mi = mask_index(row, w0, g.M);
      if (nw == 4) {
        if (p.mask_e) *reinterpret_cast<uint4*>(p.mask_e + mi) = make_uint4(we[0], we[1], we[2], we[3]);
        if (p.mask_rp) *reinterpret_cast<uint4*>(p.mask_rp + mi) = make_uint4(wr[0], wr[1], wr[2], wr[3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < nw) {
            if (p.mask_e) p.mask_e[mi + i] = we[i];
            if (p.mask_rp) p.mask_rp[mi + i] = wr[i];
          }
        }
      }
    }
    if (row < g.M) total += sum;
  }
  __device__ void finish(int, int lane) {
    if (p.tma) { slab_e.drain(lane); }
    if (!p.l1_partial) return;
    const float s = warp_sum(total);
    if (lane == 0) p.l1_partial[static_cast<size_t>(blockIdx.x) * kWarps + ew] = s;
  }
};

// ------------------------------------------------------------------------------------------------ gated dE
// acc = DIFF W_dec (unscaled dE).  With f_gate detached (gated_sae.py:39) and the aux term gradient-free (:53-54):
//   dMag' = 1[e>0] * acc                     (gradient of the magnitude pre-activation)
//   dPi'  = 1[relu_pi>0] * l1c               (only the L1 term reaches pi), l1c = lambda*C/(2F)
//   A'    = dPi' + exp(r) * dMag'            -> bf16, the single operand of the dW_gate GEMM
// The two indicator functions come from the encoder's 1-bit masks (8 B per row and warp instead of 2 x 128 B of bf16).
// Column sums over tokens, read back from two staged bf16 slabs per warp (see EpiDPreT): sum dMag' (-> db_mag) and
// sum A' (-> the rank-1 fix-up, and db_gate = sum A' - exp(r) * sum dMag').  dr_mag needs sum_t dMag' * e, which is
// sum_c W_dec[c,f] * (DIFF^T E)[c,f] -- a by-product of the dW_dec GEMM (gated_rmag_kernel), so e is never re-read.
struct EpiGatedDPre {
  struct Params {
    alignas(64) CUtensorMap tm_a;  // bf16 A' [M,N]
    const uint32_t* mask_e;        // group-major 1-bit masks of the encoder (mask_index)
    const uint32_t* mask_rp;
    const float* exp_r;            // [N]
    float* colsum_mag;             // [tiles_m, N]
    float* colsum_a;               // [tiles_m, N]
    float l1c;
    int words;
    int block_n;
    int slab_major;
  };
  static constexpr int kWarps = 8;
  static constexpr int kColVecs = 1;
  static constexpr uint32_t kSmemBytes = 2 * SlabWriter1::bytes(kWarps) + 2 * 4 * 256 * sizeof(float) + 2 * 256 * sizeof(float);
  const Params& p;
  SlabWriter1 slab_a, slab_m;  // slab_m is only staged (never stored)
  ColVecStage<1, kWarps * 32> stage;
  float* s_col;  // [2][4][256]
  float* cv_base;
  const float* cv;
  uint32_t we[4], wr[4];
  int ew, cpw, c_first;
  __device__ EpiGatedDPre(const Params& p_, uint8_t* smem, int ew_, int block_n)
      : p(p_), s_col(reinterpret_cast<float*>(smem + 2 * SlabWriter1::bytes(kWarps))),
        cv_base(reinterpret_cast<float*>(smem + 2 * SlabWriter1::bytes(kWarps)) + 2 * 4 * 256), cv(cv_base), ew(ew_),
        cpw((block_n / 32) / (kWarps / 4)), c_first((ew_ / 4) * ((block_n / 32) / (kWarps / 4))) {
    slab_a.init(smem, ew_);
    slab_m.init(smem + SlabWriter1::bytes(kWarps), ew_);
  }
  __device__ void colvec_fetch(const GemmProblem& g, const TileInfo& ti, int tid) {
    const float* const src[1] = {p.exp_r};
    stage.fetch(src, ti.n0, g.N, tid);
  }
  __device__ void colvec_commit(uint32_t parity, int tid) {
    float* dst = cv_base + parity * 256;
    stage.commit(dst, tid);
    cv = dst;
  }
  __device__ void begin_tile(const GemmProblem& g, const TileInfo& ti, int row, int, int) {
    const int w0 = (ti.n0 >> 5) + c_first;
    const int nw = max(0, min(cpw, p.words - w0));
#pragma unroll
    for (int i = 0; i < 4; ++i) { we[i] = 0; wr[i] = 0; }
    if (row < g.M && nw > 0) {
      const size_t mi = mask_index(row, w0, g.M);
      if (nw == 4) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p.mask_e + mi));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.mask_rp + mi));
        we[0] = a.x; we[1] = a.y; we[2] = a.z; we[3] = a.w;
        wr[0] = b.x; wr[1] = b.y; wr[2] = b.z; wr[3] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < nw) { we[i] = __ldg(p.mask_e + mi + i); wr[i] = __ldg(p.mask_rp + mi + i); }
      }
    }
  }
  // column sums of a finished slab for this lane's column pair (same access pattern as EpiDPreT::slab_colsum)
  __device__ __forceinline__ float2 colsum(const SlabWriter1& s, int lane) const {
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int g8 = 0; g8 < 4; ++g8) {
      __nv_bfloat162 acc;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(
            s.base + (g8 * 8 + q) * 128 + ((((lane >> 2) ^ q) << 4) | ((lane & 3) << 2)));
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w);
        acc = q == 0 ? h : __hadd2(acc, h);
      }
      const uint32_t aw = *reinterpret_cast<const uint32_t*>(&acc);
      r.x += bf16lo(aw);
      r.y += bf16hi(aw);
    }
    return r;
  }
  __device__ __forceinline__ void slab_done(const TileInfo& ti, int col_slab0, int wq, int lane) {
    __syncwarp();
    const float2 cm = colsum(slab_m, lane), ca = colsum(slab_a, lane);
    const int cc = (col_slab0 - ti.n0) + 2 * lane;  // column inside the tile
    *reinterpret_cast<float2*>(s_col + (0 * 4 + wq) * 256 + cc) = cm;
    *reinterpret_cast<float2*>(s_col + (1 * 4 + wq) * 256 + cc) = ca;
  }
  __device__ __forceinline__ void chunk(const GemmProblem& g, const TileInfo& ti, int, int col0, float (&v)[32], int wq,
                                        int lane, int ci) {
    const uint32_t be = we[ci], br = wr[ci];
    float a[32];
    const float* er = cv + (col0 - ti.n0);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] = (be & (1u << j)) ? v[j] : 0.f;                       // dMag'
      a[j] = ((br & (1u << j)) ? p.l1c : 0.f) + er[j] * v[j];     // A'
    }
    const int half = ci & 1;
    slab_a.put(half, lane, a);
    {  // dMag' is only staged for its column sums: same swizzled layout, no store
      uint8_t* rowp = slab_m.base + lane * 128;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = half * 4 + i;
        *reinterpret_cast<uint4*>(rowp + ((j ^ (lane & 7)) << 4)) = pack8_bf16(v + 8 * i);
      }
    }
    if (half == 1) {
      slab_done(ti, col0 - 32, wq, lane);
      slab_a.flush(&p.tm_a, col0 - 32, ti.m0 + wq * 32, lane, p.slab_major);
    }
  }
  __device__ void end_tile(const GemmProblem& g, const TileInfo& ti, int, int wq, int lane) {
    if (slab_a.half_pending) {  // N tail: clear the never-written second half of both slabs
      float z[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) z[j] = 0.f;
      slab_a.put(1, lane, z);
      uint8_t* rowp = slab_m.base + lane * 128;
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(rowp + (((4 + i) ^ (lane & 7)) << 4)) = make_uint4(0, 0, 0, 0);
      const int col_slab0 = ((g.N - 1) >> 6) << 6;
      slab_done(ti, col_slab0, wq, lane);
      slab_a.flush(&p.tm_a, col_slab0, ti.m0 + wq * 32, lane, p.slab_major);
    }
    // combine the four lane quarters (fixed order) and write one partial row per M tile
    epi_bar_sync(kWarps * 32);
    const int c = ew * 32 + lane;  // 256 epilogue threads, one column each
    const int col = ti.n0 + c;
    if (c < p.block_n && col < g.N) {
      const float* s0 = s_col;
      const float* s1 = s_col + 4 * 256;
      p.colsum_mag[static_cast<size_t>(ti.tile_m) * g.N + col] = (s0[c] + s0[256 + c]) + (s0[512 + c] + s0[768 + c]);
      p.colsum_a[static_cast<size_t>(ti.tile_m) * g.N + col] = (s1[c] + s1[256 + c]) + (s1[512 + c] + s1[768 + c]);
    }
    epi_bar_sync(kWarps * 32);
  }
  __device__ void finish(int, int lane) { slab_a.drain(lane); }
};

}  // namespace svb
