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
// Fused: stores of e and relu_pi, activity bits of e, sum|relu_pi| partials (sparse_loss.py:71).
struct EpiGatedEnc {
  struct Params {
    const float* dot;     // [N]
    const float* b_gate;  // [N]
    const float* b_mag;   // [N]
    const float* exp_r;   // [N]
    __nv_bfloat16* e_bf16;
    float* e_f32;
    __nv_bfloat16* rp_bf16;
    float* rp_f32;
    uint32_t* act_bits;
    float* l1_partial;  // [gridDim.x * kWarps] or null: one running sum per CTA and epilogue warp
    int hw, words;
  };
  static constexpr int kWarps = 8;
  static constexpr int kColVecs = 4;
  static constexpr uint32_t kSmemBytes = 2 * 4 * 256 * sizeof(float);
  const Params& p;
  ColVecStage<4, kWarps * 32> stage;
  float* cv_base;
  const float* cv;
  float sum, total;
  int ew;
  __device__ EpiGatedEnc(const Params& p_, uint8_t* smem, int ew_, int)
      : p(p_), cv_base(reinterpret_cast<float*>(smem)), cv(cv_base), sum(0.f), total(0.f), ew(ew_) {}
  __device__ void colvec_fetch(const GemmProblem& g, const TileInfo& ti, int tid) {
    const float* const src[4] = {p.dot, p.b_gate, p.b_mag, p.exp_r};
    stage.fetch(src, ti.n0, g.N, tid);
  }
  __device__ void colvec_commit(uint32_t parity, int tid) {
    float* dst = cv_base + parity * 4 * 256;
    stage.commit(dst, tid);
    cv = dst;
  }
  __device__ void begin_tile(const GemmProblem&, const TileInfo&, int, int, int) { sum = 0.f; }
  __device__ __forceinline__ void chunk(const GemmProblem& g, const TileInfo& ti, int row, int col0, float (&v)[32],
                                        int wq, int lane, int) {
    const int nvalid = min(32, g.N - col0);
    const bool row_ok = row < g.M;
    float rp[32];
    uint32_t word = 0;
    const float* cvt = cv + (col0 - ti.n0);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float raw = v[j] - cvt[j];
      const float pi = raw + cvt[256 + j];
      const float mag = fmaxf(cvt[768 + j] * raw + cvt[512 + j], 0.f);
      const float gate = pi > 0.f ? 1.f : (pi == 0.f ? 0.5f : 0.f);
      const float e = gate * mag;
      rp[j] = fmaxf(pi, 0.f);
      v[j] = e;
      if (j < nvalid) {
        if (e != 0.f) word |= (1u << j);
        sum += rp[j];
      }
    }
    if (!row_ok) word = 0;
    const long long off = static_cast<long long>(row) * g.N + col0;
    if (row_ok) {
      if (p.e_bf16) store_row_bf16(p.e_bf16 + off, v, nvalid);
      if (p.e_f32) store_row_f32(p.e_f32 + off, v, nvalid);
      if (p.rp_bf16) store_row_bf16(p.rp_bf16 + off, rp, nvalid);
      if (p.rp_f32) store_row_f32(p.rp_f32 + off, rp, nvalid);
    }
    if (p.act_bits) publish_activity(p.act_bits, p.words, col0 >> 5, word, row, g.M, p.hw, ti.m0 + wq * 32, lane);
  }
  __device__ void end_tile(const GemmProblem& g, const TileInfo&, int row, int, int) {
    if (row < g.M) total += sum;
  }
  __device__ void finish(int, int lane) {
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
// Column sums over tokens: sum dMag' (-> db_mag), sum dPi' (-> db_gate), sum dMag'*e (-> dr_mag).
struct EpiGatedDPre {
  struct Params {
    const __nv_bfloat16* e;
    const __nv_bfloat16* rp;
    const float* exp_r;    // [N]
    __nv_bfloat16* a_out;  // [M,N]
    float* colsum_mag;     // [tiles_m, N]
    float* colsum_pi;      // [tiles_m, N]
    float* colsum_mage;    // [tiles_m, N]
    float l1c;
    int block_n;
  };
  static constexpr int kWarps = 8;
  static constexpr int kColVecs = 1;
  static constexpr uint32_t kSmemBytes = 3 * 4 * 256 * sizeof(float) + 2 * 256 * sizeof(float);
  const Params& p;
  ColVecStage<1, kWarps * 32> stage;
  float* s_col;  // [3][4][256]
  float* cv_base;
  const float* cv;
  int ew;
  __device__ EpiGatedDPre(const Params& p_, uint8_t* smem, int ew_, int)
      : p(p_), s_col(reinterpret_cast<float*>(smem)), cv_base(reinterpret_cast<float*>(smem) + 3 * 4 * 256),
        cv(cv_base), ew(ew_) {}
  __device__ void colvec_fetch(const GemmProblem& g, const TileInfo& ti, int tid) {
    const float* const src[1] = {p.exp_r};
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
    float e[32], t[32];
    const long long off = static_cast<long long>(row) * g.N + col0;
    load_row_bf16(p.e + (row_ok ? off : 0), e, row_ok ? nvalid : 0);
    load_row_bf16(p.rp + (row_ok ? off : 0), t, row_ok ? nvalid : 0);
    float a[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float dmag = e[j] > 0.f ? v[j] : 0.f;
      const float dpi = t[j] > 0.f ? p.l1c : 0.f;
      a[j] = dpi + cv[(col0 - ti.n0) + j] * dmag;
      v[j] = dmag;
      t[j] = dpi;
      e[j] = dmag * e[j];
    }
    if (row_ok) store_row_bf16(p.a_out + off, a, nvalid);
    const int cc = (col0 - ti.n0) + lane;
    s_col[(0 * 4 + wq) * 256 + cc] = warp_colsum32(v, lane);
    s_col[(1 * 4 + wq) * 256 + cc] = warp_colsum32(t, lane);
    s_col[(2 * 4 + wq) * 256 + cc] = warp_colsum32(e, lane);
  }
  __device__ void end_tile(const GemmProblem& g, const TileInfo& ti, int, int wq, int lane) {
    epi_bar_sync(kWarps * 32);
    const int c = ew * 32 + lane;  // 256 epilogue threads, one column each
    const int col = ti.n0 + c;
    float* outs[3] = {p.colsum_mag, p.colsum_pi, p.colsum_mage};
    if (c < p.block_n && col < g.N) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float* s = s_col + q * 4 * 256;
        outs[q][static_cast<size_t>(ti.tile_m) * g.N + col] = (s[c] + s[256 + c]) + (s[512 + c] + s[768 + c]);
      }
    }
    epi_bar_sync(kWarps * 32);
  }
  __device__ void finish(int, int) {}
};

}  // namespace svb
