// SaeMLP forward and the fused SaeMLP training step (model_pipeline.py:363-432 train branch) as a stream of
// tcgen05 GEMMs with fused epilogues plus a handful of small memory-bound kernels.
//
// Data flow of one step (T tokens, C channels, F features; all big tensors bf16, token-major):
//   prep      W_enc,W_dec -> bf16 shadows; fold = b_enc - W_enc b_dec
//   pack      x (NCHW / fp32) -> X [T,C]                                  (skipped when x is already bf16 tokens)
//   G1 enc    E = relu(X W_enc^T + fold)              + activity bits + sum|E| partials
//   G2 dec    D = E W_dec^T + b_dec; DIFF = D - X     + sum DIFF^2 partials
//   stats     per-(image,channel) sums of X, D, DIFF  -> rmse/nrmse, variance explained, colsum(DIFF)
//   G3 dE     DP = 1[E>0] (DIFF W_dec + lambda*C/(2F)) + per-feature column sums (-> db_enc)
//   G4        P_wd = DIFF^T E      (split-K over tokens, fp32 partials)
//   G5        P_we = DP^T X        (split-K over tokens, fp32 partials)
//   grads     flat buffer [gW_enc | gb_enc | gW_dec | gb_dec | loss sums | max section], scaled by 2/(T_global*C)
//   -- data parallel: the caller all-reduces the flat buffer here --
//   adam      (Constrained)Adam on all four tensors; stats + activity outputs finalised
// The backward runs in units of T*C/2 (dPre' = dPre * T*C/2) so bf16 intermediates stay O(1).
#include "svb_common.cuh"
#include "fused_bwd_sm100.cuh"
#include "gemm2_sm100.cuh"

using namespace svb;

namespace {

struct SaePlan {
  long long T;
  int C, F, hw, words;
  long long n_img;
  int tiles_m, tn_f, tn_c;
  int s_wd, s_we;  // split-K slices
  int sms;
  bool zero_copy_x, bstat;  // bstat: the K <= 256 GEMMs run B-stationary (per-CTA column-sum partials)
  bool xs, es, ds;          // slab-major workspaces: X / D (xs), E / dPre' (es) and DIFF (ds), see gemm_host.cuh
  bool tok_fused;           // zero-copy token-major input (channels_last) through the fused decoder epilogue
  bool fused_dec;           // decoder epilogue writes NCHW d and the channel statistics itself (EpiDecNchw)
  bool fused_bwd;           // dE GEMM -> mask -> dW_enc GEMM in one kernel, dPre' never written (fused_bwd_sm100.cuh)
  int nt_hw;                // HW tiles of 64 positions (x statistics of the pack kernel)
  float *xpart, *dpart;
  int cs_rows;              // rows of colsum_part
  size_t zero_words;        // 32-bit words to clear at the start of a step, from act_bits on
  // workspace
  bf16 *X, *Web, *Wdb, *E, *DP, *D, *DIFF;
  float *fold, *l1_part, *sq_part, *colsum_part, *stage, *csum, *st, *chan, *var_part, *rowvar, *P_wd, *P_we, *vm,
      *nact_f, *flat;
  uint32_t *act_bits, *mask;
  // flat buffer offsets
  size_t o_gwe, o_gbe, o_gwd, o_gbd, o_sums, o_chansq, o_count, o_max;
  size_t sum_elems, max_elems;
};

constexpr int kVmChunks = 32;
// B-stationary GEMMs can let one CTA per group pull the shared A tile into L2 some steps ahead of its use
// (GemmProblem::a_prefetch).  In the stand-alone probe that takes the output-bound GEMM from 0.193 to 0.180 ms at 2
// steps (0.187 / 0.193 at 4 / 8: the lines are evicted again), but inside the step, where X / DIFF were written just
// before, it changes nothing (0.241 vs 0.236 ms), so it is off.
constexpr int kAPrefetch = 0;
// tuning(kTuneEncTwoCta) = 1 runs the encoder GEMM on SM pairs: measured 0.250 against 0.214 ms with the real epilogue
// (mask words + l1), so off by default.
bool enc_two_cta() { return tuning(kTuneEncTwoCta) != 0; }

void carve(Arena& a, SaePlan& p, const svb_acts* x, int F, bool train, int sms) {
  p.C = x->C; p.F = F; p.hw = x->hw; p.n_img = x->n_images; p.sms = sms;
  p.T = x->n_images * static_cast<long long>(x->hw);
  p.words = (F + 31) / 32;
  p.tiles_m = cdiv(p.T, kBlockM);
  p.tn_f = cdiv(F, 256);
  p.tn_c = cdiv(p.C, 256);
  p.zero_copy_x = acts_are_bf16_tokens(x);
  p.xs = false; p.es = false; p.ds = false; p.fused_dec = false; p.tok_fused = false;
  // X / D / DIFF may be slab-major with a zero-padded last slab (C % 64 != 0): size them for ceil(C / 64) slabs
  const size_t TC = static_cast<size_t>(p.T) * (cdiv(p.C, 64) * 64), TF = static_cast<size_t>(p.T) * F, FC = static_cast<size_t>(F) * p.C;
  p.X = p.zero_copy_x ? nullptr : a.take<bf16>(TC);
  p.Web = a.take<bf16>(FC);
  p.Wdb = a.take<bf16>(FC);
  p.fold = a.take<float>(F);
  p.E = a.take<bf16>(TF);
  p.D = a.take<bf16>(TC + 8 * static_cast<size_t>(p.C));   // + slack: also the channel-major [C][round8(T)] copy of d
  if (!train) return;
  p.fused_bwd = fused_bwd_enabled() && fused_bwd_supported(p.T, p.C, F, sms);
  p.DP = p.fused_bwd ? nullptr : a.take<bf16>(TF);   // the fused backward keeps dPre' on the SM
  p.DIFF = a.take<bf16>(TC);
  p.mask = a.take<uint32_t>(static_cast<size_t>(p.T) * 4 * cdiv(p.words, 4));   // group-major, see mask_index
  // cleared by the step prologue: activity bits and the per-(CTA, warp) loss partials (contiguous on purpose)
  const size_t z0 = a.off;
  p.act_bits = a.take<uint32_t>(static_cast<size_t>(p.n_img) * p.words);
  p.l1_part = a.take<float>(static_cast<size_t>(sms) * 16);   // one per CTA and epilogue warp (8 or 16 of them)
  p.sq_part = a.take<float>(static_cast<size_t>(sms) * 16);
  p.zero_words = (a.off - z0) / 4;
  p.bstat = p.C <= 256 && p.tn_f <= sms;
  p.stage = a.take<float>(static_cast<size_t>(32) * (F > p.C ? F : p.C));
  p.csum = a.take<float>(F);
  p.st = a.take<float>(stats_elems(p.n_img, p.hw, p.T, p.C));
  p.chan = a.take<float>(4 * p.C);
  p.var_part = a.take<float>(2 * cdiv(p.C, 8) + 2);
  p.rowvar = a.take<float>(p.hw == 1 ? 2 * static_cast<size_t>(p.T) : 2);
  p.s_wd = planned_splits_s(p.C, F, static_cast<int>(p.T), 0, sms);
  p.s_we = p.fused_bwd ? fused_bwd_slots(p.T, p.C, F, sms) : planned_splits<256>(F, p.C, static_cast<int>(p.T), 0);
  // per-feature column sums of dPre': one row per split of the dW_enc GEMM (EpiPartialOnes), or two (token halves) per
  // slot of the fused backward
  p.cs_rows = p.fused_bwd ? 2 * p.s_we : p.s_we;
  p.colsum_part = a.take<float>(static_cast<size_t>(p.cs_rows) * F);
  p.P_wd = a.take<float>(static_cast<size_t>(p.s_wd) * FC);
  p.P_we = a.take<float>(static_cast<size_t>(p.s_we) * FC);
  p.vm = a.take<float>(static_cast<size_t>(kVmChunks) * p.C);
  p.nt_hw = cdiv(p.hw, 64);
  p.xpart = a.take<float>(static_cast<size_t>(p.n_img) * p.nt_hw * 4 * p.C);
  p.dpart = a.take<float>(static_cast<size_t>(p.tiles_m) * 4 * 2 * 3 * p.C);
  p.nact_f = a.take<float>(p.n_img);
  p.o_gwe = 0; p.o_gbe = FC; p.o_gwd = FC + F; p.o_gbd = 2 * FC + F;
  p.o_sums = 2 * FC + F + p.C;
  p.o_chansq = p.o_sums + 8;
  p.o_count = p.o_chansq + p.C;
  p.sum_elems = p.o_count + F;
  p.o_max = p.sum_elems;
  p.max_elems = 2 * static_cast<size_t>(p.C);
  p.flat = a.take<float>(p.sum_elems + p.max_elems);
}

int plan(svb_handle* h, SaePlan& p, const svb_acts* x, int F, bool train) {
  Arena dry;
  dry.dry = true;
  carve(dry, p, x, F, train, h->sms);
  SVB_TRY(ensure_arena(h, dry.off));
  h->arena.off = 0;
  h->arena.dry = false;
  carve(h->arena, p, x, F, train, h->sms);
  if (train) p.flat = comm_flat_or(h, p.flat, p.sum_elems + p.max_elems);  // data parallel: exchange buffer in peer memory
  return 0;
}

int check_params(const svb_acts* x, const svb_sae_params* p) {
  SVB_TRY(check_acts(x));
  if (!p || !p->w_enc || !p->b_enc || !p->w_dec || !p->b_dec) return fail(SVB_ERR_BAD_ARG, "null SAE parameter");
  if (p->F <= 0 || p->F % 8) return fail(SVB_ERR_UNSUPPORTED, "hidden_size F=%d must be a positive multiple of 8", p->F);
  return 0;
}

int run_prep(cudaStream_t st, const SaePlan& pl, const svb_sae_params* p, bool train) {
  PrepArgs a{};
  a.w_enc = p->w_enc; a.b_enc = p->b_enc; a.b_dec = p->b_dec; a.w_enc_bf16 = pl.Web; a.fold = pl.fold;
  a.w_dec = p->w_dec; a.w_dec_bf16 = pl.Wdb;
  a.zero = train ? pl.act_bits : nullptr; a.n_zero = train ? pl.zero_words : 0;
  a.F = pl.F; a.C = pl.C;
  return run_prep_step(st, a);
}

}  // namespace

extern "C" int svb_sae_forward(svb_handle* h, void* stream, const svb_acts* x, const svb_sae_params* p,
                               const svb_sae_forward_out* out) {
  if (!h || !out) return fail(SVB_ERR_BAD_ARG, "null handle/out");
  SVB_ON_DEVICE(h);
  SVB_TRY(check_params(x, p));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SaePlan pl;
  SVB_TRY(plan(h, pl, x, p->F, false));
  h->gradbuf = nullptr;
  const bf16* X = pl.zero_copy_x ? static_cast<const bf16*>(x->x) : pl.X;
  if (!pl.zero_copy_x) SVB_TRY(pack_acts(st, x, pl.X));
  SVB_TRY(run_prep(st, pl, p, false));
  const int T = static_cast<int>(pl.T);
  EpiEncApi::Params e1{};
  e1.bias = pl.fold;
  e1.e_bf16 = (out->enc && out->enc_dtype == SVB_BF16) ? static_cast<bf16*>(out->enc) : pl.E;
  e1.e_f32 = (out->enc && out->enc_dtype == SVB_F32) ? static_cast<float*>(out->enc) : nullptr;
  e1.pre_f32 = out->pre;
  e1.words = pl.words;
  if (make_store_tmap_bf16_chunk(&e1.tm_e, e1.e_bf16, T, pl.F, pl.F)) return fail(SVB_ERR_TMAP, "tensor map for enc output");
  SVB_GEMM((launch_gemm_s<false, false, EpiEncApi>(st, X, pl.C, pl.Web, pl.C, T, pl.F, pl.C, 1, e1)), "enc");
  if (out->dec) {
    EpiDec::Params e2{};
    e2.bias = p->b_dec;
    e2.d_bf16 = out->dec_dtype == SVB_BF16 ? static_cast<bf16*>(out->dec) : nullptr;
    e2.d_f32 = out->dec_dtype == SVB_F32 ? static_cast<float*>(out->dec) : nullptr;
    if (e2.d_bf16 && make_store_tmap_bf16(&e2.tm_d, e2.d_bf16, T, pl.C, pl.C)) return fail(SVB_ERR_TMAP, "tensor map for dec output");
    SVB_GEMM((launch_gemm_s<false, false, EpiDec>(st, e1.e_bf16, pl.F, pl.Wdb, pl.F, T, pl.C, pl.F, 1, e2)), "dec");
  }
  return 0;
}

extern "C" int svb_sae_step_grads(svb_handle* h, void* stream, const svb_acts* x, const svb_sae_params* p,
                                  float lambda_sparse, int64_t global_tokens, const svb_train_out* out) {
  if (!h) return fail(SVB_ERR_BAD_ARG, "null handle");
  SVB_ON_DEVICE(h);
  SVB_TRY(check_params(x, p));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SaePlan pl;
  SVB_TRY(plan(h, pl, x, p->F, true));
  const int T = static_cast<int>(pl.T), C = pl.C, F = pl.F;
  const double Tg = global_tokens > 0 ? static_cast<double>(global_tokens) : static_cast<double>(pl.T);
  const bf16* X = pl.zero_copy_x ? static_cast<const bf16*>(x->x) : pl.X;
  // Slab-major workspaces (every 128 x 64 operand tile / 32 x 64 epilogue slab is one contiguous block):
  // E and dPre' whenever F % 64 == 0; X, D, DIFF when we pack X ourselves from NCHW and the fused post-decoder pass runs.
  pl.es = F % 64 == 0;
  const bool slab_ok = !pl.zero_copy_x && post_dec_fusable(x, out ? out->dec_out : nullptr, out ? out->dec_layout : SVB_NCHW);
  // Fused decoder epilogue: needs the slab-major path and >= 32 tokens per image (a warp's 32 tokens then touch at
  // most two images).  bf16 d goes back to the caller's NCHW tensor by TMA when that can address it (16-byte row pitch:
  // HW % 8 == 0); every other output (14x14 / 7x7 maps, fp32) leaves as a channel-major bf16 TMA copy that a small
  // vectorised kernel scatters / widens into the NCHW tensor beside the dE GEMM -- measured faster than 8- / 16-byte
  // stores from the epilogue warps (same box: mixed4a 0.63 -> 0.60 ms bf16, mixed3b fp32 2.00 -> 1.95 ms).
  void* dec_out = out ? out->dec_out : nullptr;
  // Token-major bf16 activations of a conv layer (a channels_last base model: [B,C,H,W] whose memory IS the [(b h w), C]
  // token matrix) are read in place by the encoder / dW_enc GEMMs and the fused decoder epilogue, and d goes back in
  // the same format through one TMA store per chunk: no pack pass, no layout copy anywhere (models/sae_mlp.py:44).
  // Their x statistics come from a read-only kernel beside the encoder GEMM.
  pl.tok_fused = pl.zero_copy_x && x->layout == SVB_TOKENS && pl.hw >= 32 && x->n_images <= 65535 &&
                 (!dec_out || (out->dec_layout == SVB_TOKENS && out->dec_dtype == SVB_BF16 &&
                               (reinterpret_cast<uintptr_t>(dec_out) & 15) == 0));
  pl.fused_dec = (slab_ok && pl.hw >= 32) || pl.tok_fused;
  // C % 64 != 0 (mixed3b: 480, mixed4d: 528): only the fused epilogue keeps the last slab's padding columns zero
  pl.xs = slab_ok && (C % 64 == 0 || pl.fused_dec);
  pl.ds = pl.xs || pl.tok_fused;
  int out_kind = 0;
  if (dec_out && pl.tok_fused) out_kind = 2;
  else if (dec_out)
    out_kind = (out->dec_dtype == SVB_BF16 && pl.hw % 8 == 0 && (reinterpret_cast<uintptr_t>(dec_out) & 15) == 0) ? 1 : 4;
  const long long ld_t = (pl.T + 7) & ~7LL;   // row pitch of the channel-major copy (out_kind 4)
  h->step_flags = pl.fused_bwd ? SVB_STEP_FUSED_BWD : 0;
  prof_begin_step(h);
  prof_mark(h, st, 0);
  // weight prologue on the side stream, next to the activation pack
  SVB_TRY(side_fork(h, st));
  SVB_TRY(run_prep(h->side, pl, p, true));
  if (!pl.zero_copy_x) SVB_TRY(pack_acts(st, x, pl.X, pl.xs, pl.fused_dec ? pl.xpart : nullptr));
  // zero-copy tokens: only their statistics are needed (one read-only pass, where the pack pass would be).  Measured:
  // hidden on the side stream beside the encoder GEMM it gets the 4 SMs the 144-CTA GEMM leaves free, runs for the
  // whole encoder AND decoder phase and costs the decoder 0.03 ms; here it is 0.02 ms beside the weight prologue.
  if (pl.tok_fused) {
    launch_x_stats_tokens(st, X, pl.xpart, C, pl.hw, pl.nt_hw, pl.n_img);
    SVB_LAUNCH_CHECK("x_stats_tokens");
  }
  SVB_TRY(side_join(h, st));

  prof_mark(h, st, 1);
  // G1 encoder
  EpiEnc::Params e1{};
  e1.bias = pl.fold; e1.e_bf16 = pl.E; e1.l1_partial = pl.l1_part;
  e1.mask_words = pl.mask;   // the per-image activity bits are derived from the masks below, not in the epilogue
  e1.words = pl.words; e1.e_slab = pl.es;
  if (pl.es ? make_store_tmap_bf16_slab32(&e1.tm_e, pl.E, T, F) : make_store_tmap_bf16_chunk(&e1.tm_e, pl.E, T, F, F))
    return fail(SVB_ERR_TMAP, "tensor map for E");
  if (pl.bstat && enc_two_cta() && gemm2_groups(T, F, pl.sms) >= 1) {
    // SM pairs (cta_group::2): each CTA keeps half of the resident weight tile and reads 8 KB instead of 12 KB of
    // operands per MMA step from shared memory, which this output-bound GEMM shares with its own TMA stores
    EpiEnc4::Params e14{};
    e14.tm_e = e1.tm_e; e14.bias = e1.bias; e14.e_bf16 = e1.e_bf16; e14.l1_partial = e1.l1_partial;
    e14.mask_words = e1.mask_words; e14.words = e1.words; e14.e_slab = e1.e_slab;
    SVB_GEMM((launch_gemm2_bstat<false, EpiEnc4>(st, X, C, pl.Web, C, T, F, C, e14, pl.xs, pl.sms)), "enc (two-CTA B-stationary)");
  } else if (pl.bstat && tuning(kTuneEnc16) != 0) {
    // 16 epilogue warps (off by default): the epilogue (~260 instructions per 32 x 32 chunk) bounds this GEMM and two
    // warps per scheduler issue only half of the cycles, but 32 KB of staging leave four operand stages instead of five
    // and the variant measured 0.221 against 0.212 ms
    EpiEnc16::Params e16{};
    e16.tm_e = e1.tm_e; e16.bias = e1.bias; e16.e_bf16 = e1.e_bf16; e16.l1_partial = e1.l1_partial;
    e16.mask_words = e1.mask_words; e16.words = e1.words; e16.e_slab = e1.e_slab;
    SVB_GEMM((launch_gemm<256, false, false, EpiEnc16, true>(st, X, C, pl.Web, C, T, F, C, 1, e16, nullptr, 0, 0, pl.xs, false, kAPrefetch)), "enc (B-stationary, 16 epilogue warps)");
  } else if (pl.bstat) {
    SVB_GEMM((launch_gemm<256, false, false, EpiEnc, true>(st, X, C, pl.Web, C, T, F, C, 1, e1, nullptr, 0, 0, pl.xs, false, kAPrefetch)), "enc (B-stationary)");
  } else {
    SVB_GEMM((launch_gemm_s<false, false, EpiEnc>(st, X, C, pl.Web, C, T, F, C, 1, e1, nullptr, 0, 0, pl.xs, false)), "enc");
  }
  prof_mark(h, st, 2);
  // per-image activity bits from the masks: side stream, beside the decoder GEMM (joined before the assembly)
  SVB_TRY(side_fork(h, st));
  (mask_to_activity_kernel<<<dim3(static_cast<unsigned>(pl.n_img), cdiv(pl.words, 4)), 128, 0, h->side>>>(
      pl.mask, pl.act_bits, pl.T, pl.hw, pl.words), svb::count_launch());
  SVB_LAUNCH_CHECK("mask_to_activity");
  // G2 decoder
  if (pl.fused_dec) {
    EpiDecNchw::Params e2{};
    e2.bias = p->b_dec; e2.x = X; e2.sq_partial = pl.sq_part; e2.part = pl.dpart; e2.hw = pl.hw;
    e2.out = dec_out; e2.out_kind = out_kind; e2.x_slab = pl.xs ? 1 : 0; e2.tok = pl.tok_fused ? 1 : 0;
    if (make_store_tmap_bf16_slab32(&e2.tm_diff, pl.DIFF, T, C)) return fail(SVB_ERR_TMAP, "tensor map for DIFF");
    if (out_kind == 2 && make_store_tmap_bf16_chunk(&e2.tm_out, dec_out, T, C, C)) return fail(SVB_ERR_TMAP, "tensor map for the token-major output");
    if (out_kind == 1 && make_tmap_nchw_bf16(&e2.tm_out, dec_out, pl.n_img, C, pl.hw)) return fail(SVB_ERR_TMAP, "tensor map for the NCHW output");
    if (out_kind == 4 && make_store_tmap_bf16_cmajor(&e2.tm_out, pl.D, C, pl.T, ld_t)) return fail(SVB_ERR_TMAP, "tensor map for the channel-major output");
    // the encoder wrote E from the first token tile to the last, so its newest ~100 MB are still in L2: walk the
    // token tiles backwards and the decoder's first reads are hits (-3.5 us of 210)
    SVB_GEMM((launch_gemm_s<false, false, EpiDecNchw>(st, pl.E, F, pl.Wdb, F, T, C, F, 1, e2, nullptr, 0, 0, pl.es, false, 0, /*reverse_m=*/true)), "dec (fused NCHW)");
    prof_mark(h, st, 3);
    // the statistics folds only feed the tail of the step: side stream, beside the dE GEMM
    SVB_TRY(side_fork(h, st));
    if (out_kind == 4) SVB_TRY(run_cmajor_to_nchw(h->side, pl.D, dec_out, out->dec_dtype, C, pl.hw, pl.T, ld_t));
    (dec_stats_image_kernel<<<dim3(static_cast<unsigned>(pl.n_img), cdiv(C, 64)), 256, 0, h->side>>>(pl.dpart, pl.xpart, pl.st, C, pl.hw, pl.nt_hw, pl.T), svb::count_launch());
    (dec_stats_channel_kernel<<<cdiv(C, 32), 1024, 0, h->side>>>(pl.st, pl.chan, pl.var_part, static_cast<int>(pl.n_img), C), svb::count_launch());
    SVB_LAUNCH_CHECK("decoder statistics");
  } else {
    EpiDec::Params e2{};
    e2.bias = p->b_dec; e2.x = X; e2.d_bf16 = pl.D; e2.diff_bf16 = pl.DIFF; e2.sq_partial = pl.sq_part;
    e2.out_slab = pl.xs; e2.x_slab = pl.xs;
    if (pl.xs ? (make_store_tmap_bf16_slab(&e2.tm_d, pl.D, T, C) || make_store_tmap_bf16_slab(&e2.tm_diff, pl.DIFF, T, C))
              : (make_store_tmap_bf16(&e2.tm_d, pl.D, T, C, C) || make_store_tmap_bf16(&e2.tm_diff, pl.DIFF, T, C, C)))
      return fail(SVB_ERR_TMAP, "tensor maps for D / DIFF");
    SVB_GEMM((launch_gemm_s<false, false, EpiDec>(st, pl.E, F, pl.Wdb, F, T, C, F, 1, e2, nullptr, 0, 0, pl.es, false)), "dec");
    prof_mark(h, st, 3);
    // channel statistics + the decoder output handed back to the model (model_pipeline.py:425,432), one pass
    SVB_TRY(run_post_dec(st, x, X, pl.D, pl.T, dec_out, out ? out->dec_dtype : SVB_BF16,
                         out ? out->dec_layout : SVB_NCHW, pl.st, pl.chan, pl.var_part, pl.rowvar, pl.xs));
  }
  prof_mark(h, st, 4);
  const float l1c = static_cast<float>(static_cast<double>(lambda_sparse) * C / (2.0 * F));
  // Weight gradients, split-K over tokens.  The encoder side goes first: with its assembly done, the leading part of
  // the flat buffer [gW_enc | gb_enc] is final and can be all-reduced while the decoder weight-gradient GEMM runs.
  const size_t FC = static_cast<size_t>(F) * C;
  const float s = static_cast<float>(2.0 / (Tg * C));
  float* flat = pl.flat;
  if (pl.fused_bwd) {
    // G3 + G5 in one kernel: dPre' = 1[E>0] (DIFF W_dec + lambda*C/(2F)) goes from TMEM through shared memory straight
    // into the dW_enc MMA; P_we holds one partial per token slot, colsum_part two rows per slot
    SVB_GEMM(launch_fused_bwd(st, pl.Wdb, pl.DIFF, pl.ds, C, X, pl.xs, C, pl.mask, T, C, F, l1c, pl.P_we, pl.colsum_part, pl.sms),
             "fused dE -> dW_enc");
    prof_mark(h, st, 5);
  } else {
    // G3 dE -> dPre'
    EpiDPreNoSum::Params e3{};
    e3.mask_words = pl.mask; e3.words = pl.words; e3.colsum_partial = nullptr; e3.l1c = l1c; e3.out_slab = pl.es;
    if (pl.es ? make_store_tmap_bf16_slab32(&e3.tm_dpre, pl.DP, T, F) : make_store_tmap_bf16_chunk(&e3.tm_dpre, pl.DP, T, F, F))
      return fail(SVB_ERR_TMAP, "tensor map for dPre");
    if (pl.bstat) {
      SVB_GEMM((launch_gemm<256, false, true, EpiDPreNoSum, true>(st, pl.DIFF, C, pl.Wdb, F, T, F, C, 1, e3, nullptr, 0, 0, pl.ds, false, kAPrefetch)), "dE (B-stationary)");
    } else {
      SVB_GEMM((launch_gemm_s<false, true, EpiDPreNoSum>(st, pl.DIFF, C, pl.Wdb, F, T, F, C, 1, e3, nullptr, 0, 0, pl.ds, false)), "dE");
    }
    prof_mark(h, st, 5);
    // (its extra ones column yields the per-feature column sums of dPre' that db_enc and the rank-1 fix-up need)
    EpiPartialOnes::Params e5{pl.P_we, C, static_cast<long long>(FC), pl.colsum_part};
    SVB_GEMM((launch_gemm<256, true, true, EpiPartialOnes>(st, pl.DP, F, X, C, F, C, T, 0, e5, nullptr, 0, 0, pl.es, pl.xs)), "dW_enc");
  }
  // column-sum reduction + encoder-side assembly on the side stream, beside the dW_dec GEMM
  SVB_TRY(side_fork(h, st));
  SVB_TRY(reduce_rows(h->side, pl.colsum_part, pl.cs_rows, F, 1.f, pl.stage, pl.csum));
  AssembleArgs aa{};
  aa.P_wd = pl.P_wd; aa.g_wdec = flat + pl.o_gwd; aa.s_wd = pl.s_wd;
  aa.P_we = pl.P_we; aa.g_wenc = flat + pl.o_gwe; aa.s_we = pl.s_we;
  aa.csum = pl.csum; aa.b_dec = p->b_dec; aa.g_benc = flat + pl.o_gbe;
  aa.w_enc_bf16 = pl.Web; aa.vm = pl.vm; aa.vm_chunks = kVmChunks;
  aa.act_bits = pl.act_bits; aa.count = flat + pl.o_count; aa.n_active = out ? out->activity.n_active : nullptr;
  aa.nact_f = pl.nact_f; aa.n_img = static_cast<int>(pl.n_img); aa.words = pl.words;
  aa.F = F; aa.C = C; aa.s = s;
  SVB_TRY(run_assemble(h->side, aa, 1));
  SVB_TRY(release_comm_stream(h, h->side));
  // the one-block tail (decoder-bias gradient, loss sums, per-channel statistics) needs nothing from the dW_dec GEMM
  TailArgs ta{};
  ta.chan = pl.chan; ta.vm = pl.vm; ta.vm_chunks = kVmChunks; ta.g_bdec = flat + pl.o_gbd; ta.s = s;
  ta.sq_part = pl.sq_part; ta.n_sq = pl.sms * 16;
  ta.l1_part = pl.l1_part; ta.n_l1 = pl.sms * 16;
  ta.nact_f = pl.nact_f; ta.n_img = static_cast<int>(pl.n_img);
  ta.var_part = pl.var_part; ta.n_var_part = pl.fused_dec ? cdiv(C, 32) : cdiv(C, 8);
  ta.rowvar = pl.rowvar; ta.n_rows = pl.hw == 1 ? pl.T : 0;
  ta.flat = flat; ta.o_sums = pl.o_sums; ta.o_chansq = pl.o_chansq; ta.o_max = pl.o_max; ta.C = C;
  (grads_tail_kernel<<<1, 1024, 0, h->side>>>(ta), svb::count_launch());

  h->early_elems = h->comm ? static_cast<int64_t>(pl.o_gwd) : 0;
  prof_mark(h, st, 6);
  EpiPartial::Params e4{pl.P_wd, F, static_cast<long long>(FC)};
  SVB_GEMM((launch_gemm_s<true, true, EpiPartial>(st, pl.DIFF, C, pl.E, F, C, F, T, 0, e4, nullptr, 0, 0, pl.ds, pl.es)), "dW_dec");
  prof_mark(h, st, 7);
  // rest of the gradient assembly: the decoder weight gradient, after everything forked above has joined
  SVB_TRY(side_join(h, st));
  SVB_TRY(run_assemble(st, aa, 2));
  SVB_LAUNCH_CHECK("grad assembly");
  prof_mark(h, st, 8);
  h->gradbuf = flat;
  h->sum_elems = static_cast<int64_t>(pl.sum_elems);
  h->max_elems = static_cast<int64_t>(pl.max_elems);
  return 0;
}

extern "C" int svb_sae_step_apply(svb_handle* h, void* stream, const svb_acts* x, const svb_sae_params* p,
                                  const svb_adam_state* adam, const svb_opt_config* opt, float lambda_sparse,
                                  int32_t expansion_factor, int64_t global_tokens, int64_t global_images,
                                  const svb_train_out* out) {
  if (!h || !adam || !opt) return fail(SVB_ERR_BAD_ARG, "null handle/adam/opt");
  SVB_ON_DEVICE(h);
  SVB_TRY(check_params(x, p));
  for (int i = 0; i < 4; ++i)
    if (!adam->m[i] || !adam->v[i]) return fail(SVB_ERR_BAD_ARG, "null Adam state tensor %d", i);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SaePlan pl;
  SVB_TRY(plan(h, pl, x, p->F, true));
  if (pl.flat != h->gradbuf) return fail(SVB_ERR_BAD_ARG, "svb_sae_step_apply called without a matching svb_sae_step_grads");
  const int C = pl.C, F = pl.F;
  const size_t FC = static_cast<size_t>(F) * C;
  float* flat = pl.flat;
  AdamCoef k;
  SVB_TRY(adam_coef_for(h, st, opt, &k));
  AdamSeg segs[4];
  int ns = 0;
  segs[ns++] = AdamSeg{p->w_enc, flat + pl.o_gwe, adam->m[0], adam->v[0], FC};
  segs[ns++] = AdamSeg{p->b_enc, flat + pl.o_gbe, adam->m[1], adam->v[1], static_cast<unsigned long long>(F)};
  segs[ns++] = AdamSeg{p->b_dec, flat + pl.o_gbd, adam->m[3], adam->v[3], static_cast<unsigned long long>(C)};
  if (opt->optimizer != SVB_CONSTRAINED_ADAM) segs[ns++] = AdamSeg{p->w_dec, flat + pl.o_gwd, adam->m[2], adam->v[2], FC};
  // The three pieces of the tail are independent of each other (different tensors; the finalise block only reads the
  // flat buffer): the constrained-Adam kernel stays on the caller's stream, plain Adam and the finalise block run
  // beside it on the side stream (-0.012 ms of launch-to-launch latency at cfg2).
  const bool cadam = opt->optimizer == SVB_CONSTRAINED_ADAM;
  SVB_TRY(side_fork(h, st));
  cudaStream_t s2 = h->side;
  SVB_TRY(run_adam_multi(cadam ? s2 : st, segs, ns, k));
  if (cadam) launch_cadam(st, p->w_dec, flat + pl.o_gwd, adam->m[2], adam->v[2], C, F, k);
  SVB_LAUNCH_CHECK("adam");
  if (out && (out->stats || out->activity.dead || out->activity.freq)) {
    FinalizeArgs fa{};
    fa.flat = flat; fa.o_sums = pl.o_sums; fa.o_chansq = pl.o_chansq; fa.o_max = pl.o_max; fa.o_count = pl.o_count;
    fa.C = C; fa.F = F; fa.expansion = expansion_factor;
    fa.T_g = static_cast<float>(global_tokens > 0 ? global_tokens : pl.T);
    fa.B_g = static_cast<float>(global_images > 0 ? global_images : pl.n_img);
    fa.lambda = lambda_sparse;
    fa.stats = out->stats; fa.dead = out->activity.dead; fa.freq = out->activity.freq;
    (step_finalize_kernel<<<1, 1024, 0, s2>>>(fa), svb::count_launch());
    SVB_LAUNCH_CHECK("finalize");
  }
  SVB_TRY(side_join(h, st));
  prof_mark(h, st, 9);
  return 0;
}

extern "C" int svb_sae_train_step(svb_handle* h, void* stream, const svb_acts* x, const svb_sae_params* p,
                                  const svb_adam_state* adam, const svb_opt_config* opt, float lambda_sparse,
                                  int32_t expansion_factor, const svb_train_out* out) {
  SVB_TRY(svb_sae_step_grads(h, stream, x, p, lambda_sparse, 0, out));
  return svb_sae_step_apply(h, stream, x, p, adam, opt, lambda_sparse, expansion_factor, 0, 0, out);
}

extern "C" int svb_sae_grad_buffer(svb_handle* h, float** buf, int64_t* sum_elems, int64_t* max_elems) {
  if (!h || !h->gradbuf) return fail(SVB_ERR_BAD_ARG, "no gradient buffer: call svb_*_step_grads first");
  if (buf) *buf = h->gradbuf;
  if (sum_elems) *sum_elems = h->sum_elems;
  if (max_elems) *max_elems = h->max_elems;
  return 0;
}
