// GatedSae forward and fused training step (models/gated_sae.py:28-56, losses/sparse_loss.py:68-76,
// utils.py:2455-2473, model_pipeline.py:380-388).  Six GEMMs: shared gate/magnitude encoder, decoder, frozen-decoder
// "via gate" pass (loss value only — it is computed under no_grad in the reference and carries no gradient), dE,
// dW_dec and a single dW_gate GEMM whose operand A' = dPi + exp(r_mag) * dMag merges both sub-layer gradients.
#include "svb_common.cuh"
#include "gemm2_sm100.cuh"
#include "epilogues_gated.cuh"

using namespace svb;

namespace {

struct GatedPlan {
  long long T, n_img;
  int C, F, hw, words, tiles_m, tn_f, tn_c, s_wd, s_wg, sms;
  size_t zero_words;  // 32-bit words cleared by the step prologue, from act_bits on
  bool zero_copy_x;
  bf16 *X, *Wgb, *Wdb, *E, *RP, *A, *D, *DIFF;
  float *dot, *exp_r, *l1_part, *sq_part, *aux_part, *cs_mag, *cs_a, *stage, *csum_mag, *csum_a, *st, *chan,
      *var_part, *rowvar, *P_wd, *P_wg, *vm, *nact_f, *flat;
  uint32_t *act_bits, *mask_e, *mask_rp, *cnt_part;
  int cnt_chunks;
  bool es;  // slab-major E / relu_pi / A' (F % 64 == 0), see gemm_host.cuh
  bool xs, fused_dec;   // slab-major X / DIFF and the fused NCHW decoder epilogue, as in svb_sae.cu
  int nt_hw;
  float *xpart, *dpart;
  size_t o_gwg, o_gbg, o_gbm, o_gr, o_gwd, o_gbd, o_sums, o_chansq, o_count, o_max, sum_elems, max_elems;
};

constexpr int kVmChunks = 32;
constexpr int kCountRows = 128;  // rows per block of mask_colcount_kernel (byte-lane counters: must stay <= 255)

void carve(Arena& a, GatedPlan& p, const svb_acts* x, int F, bool train, int sms) {
  p.C = x->C; p.F = F; p.hw = x->hw; p.n_img = x->n_images; p.sms = sms;
  p.T = x->n_images * static_cast<long long>(x->hw);
  p.words = (F + 31) / 32;
  p.tiles_m = cdiv(p.T, kBlockM);
  p.tn_f = cdiv(F, 256);
  p.tn_c = cdiv(p.C, 256);
  p.zero_copy_x = acts_are_bf16_tokens(x);
  p.es = F % 64 == 0;
  p.xs = false; p.fused_dec = false;
  // X / D / DIFF may be slab-major with a zero-padded last slab (C % 64 != 0): size them for ceil(C / 64) slabs
  const size_t TC = static_cast<size_t>(p.T) * (cdiv(p.C, 64) * 64) + 8 * static_cast<size_t>(p.C), TF = static_cast<size_t>(p.T) * F, FC = static_cast<size_t>(F) * p.C;
  p.X = p.zero_copy_x ? nullptr : a.take<bf16>(TC);
  p.Wgb = a.take<bf16>(FC);
  p.Wdb = a.take<bf16>(FC);
  p.dot = a.take<float>(F);
  p.exp_r = a.take<float>(F);
  p.E = a.take<bf16>(TF);
  p.RP = a.take<bf16>(TF);
  p.D = a.take<bf16>(TC);
  if (!train) return;
  p.A = a.take<bf16>(TF);
  p.DIFF = a.take<bf16>(TC);
  const size_t z0 = a.off;
  p.act_bits = a.take<uint32_t>(static_cast<size_t>(p.n_img) * p.words);
  p.l1_part = a.take<float>(static_cast<size_t>(sms) * 8);
  p.sq_part = a.take<float>(static_cast<size_t>(sms) * 8);
  p.aux_part = a.take<float>(static_cast<size_t>(sms) * 8);
  p.zero_words = (a.off - z0) / 4;
  p.mask_e = a.take<uint32_t>(static_cast<size_t>(p.T) * 4 * cdiv(p.words, 4));    // group-major, see mask_index
  p.mask_rp = a.take<uint32_t>(static_cast<size_t>(p.T) * 4 * cdiv(p.words, 4));
  p.cnt_chunks = cdiv(p.T, kCountRows);
  p.cnt_part = a.take<uint32_t>(static_cast<size_t>(p.cnt_chunks) * p.words * 32);
  p.cs_mag = a.take<float>(static_cast<size_t>(p.tiles_m) * F);
  p.cs_a = a.take<float>(static_cast<size_t>(p.tiles_m) * F);
  p.stage = a.take<float>(static_cast<size_t>(32) * (F > p.C ? F : p.C));
  p.csum_mag = a.take<float>(F);
  p.csum_a = a.take<float>(F);
  p.st = a.take<float>(stats_elems(p.n_img, p.hw, p.T, p.C));
  p.chan = a.take<float>(4 * p.C);
  p.var_part = a.take<float>(2 * cdiv(p.C, 8) + 2);
  p.rowvar = a.take<float>(p.hw == 1 ? 2 * static_cast<size_t>(p.T) : 2);
  p.s_wd = planned_splits_s(p.C, F, static_cast<int>(p.T), 0, sms);
  p.s_wg = planned_splits_s(F, p.C, static_cast<int>(p.T), 0, sms);
  p.P_wd = a.take<float>(static_cast<size_t>(p.s_wd) * FC);
  p.P_wg = a.take<float>(static_cast<size_t>(p.s_wg) * FC);
  p.vm = a.take<float>(static_cast<size_t>(kVmChunks) * p.C);
  p.nact_f = a.take<float>(p.n_img);
  p.nt_hw = cdiv(p.hw, 64);
  p.xpart = a.take<float>(static_cast<size_t>(p.n_img) * p.nt_hw * 4 * p.C);
  p.dpart = a.take<float>(static_cast<size_t>(p.tiles_m) * 4 * 2 * 3 * p.C);
  p.o_gwg = 0; p.o_gbg = FC; p.o_gbm = FC + F; p.o_gr = FC + 2 * static_cast<size_t>(F);
  p.o_gwd = FC + 3 * static_cast<size_t>(F);
  p.o_gbd = 2 * FC + 3 * static_cast<size_t>(F);
  p.o_sums = p.o_gbd + p.C;
  p.o_chansq = p.o_sums + 8;
  p.o_count = p.o_chansq + p.C;
  p.sum_elems = p.o_count + F;
  p.o_max = p.sum_elems;
  p.max_elems = 2 * static_cast<size_t>(p.C);
  p.flat = a.take<float>(p.sum_elems + p.max_elems);
}

int plan(svb_handle* h, GatedPlan& p, const svb_acts* x, int F, bool train) {
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

int check_params(const svb_acts* x, const svb_gated_params* p) {
  SVB_TRY(check_acts(x));
  if (!p || !p->w_gate || !p->b_gate || !p->b_mag || !p->r_mag || !p->w_dec || !p->b_dec)
    return fail(SVB_ERR_BAD_ARG, "null Gated-SAE parameter");
  if (p->F <= 0 || p->F % 8) return fail(SVB_ERR_UNSUPPORTED, "hidden_size F=%d must be a positive multiple of 8", p->F);
  return 0;
}

// Number of tokens with relu_pi > 0 per feature, from the encoder's group-major 1-bit mask (exact integers; db_gate
// = l1c * count must not be formed as the difference of two bf16-staged column sums).  A thread owns one 32-feature
// word column over a chunk of 128 rows and counts 4 bit positions per add in byte lanes (bits k, k+8, k+16, k+24).
// grid (ceil(words/128), row chunks), 128 threads;  part[chunk][f] uint32.
__global__ void __launch_bounds__(128)
mask_colcount_kernel(const uint32_t* __restrict__ mask, long long T, int words, uint32_t* __restrict__ part) {
  const int w = blockIdx.x * 128 + threadIdx.x;
  if (w >= words) return;
  const long long r0 = static_cast<long long>(blockIdx.y) * kCountRows, r1 = min(T, r0 + kCountRows);
  uint32_t lanes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long r = r0; r < r1; ++r) {
    const uint32_t v = __ldg(mask + mask_index(r, w & ~3, T) + (w & 3));
#pragma unroll
    for (int k = 0; k < 8; ++k) lanes[k] += (v >> k) & 0x01010101u;
  }
  uint32_t* o = part + (static_cast<size_t>(blockIdx.y) * words + w) * 32;
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int q = 0; q < 4; ++q) o[k + 8 * q] = (lanes[k] >> (8 * q)) & 0xFFu;
}
// per-feature vector gradients:  gb_gate = s * l1c * count(relu_pi > 0);  gb_mag = s * csum_mag
__global__ void gated_vec_grads_kernel(const float* __restrict__ cs_mag, const uint32_t* __restrict__ cnt_part,
                                       int chunks, int pitch /* words * 32 */, float l1c, float s, int F,
                                       float* __restrict__ g_bgate, float* __restrict__ g_bmag) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  uint32_t n = 0;
  for (int k = 0; k < chunks; ++k) n += cnt_part[static_cast<size_t>(k) * pitch + f];
  g_bgate[f] = s * l1c * static_cast<float>(n);
  g_bmag[f] = s * cs_mag[f];
}
// gr_mag = s * (sum_t dMag'*e - b_mag*sum_t dMag')      (d mag_pre / d r = exp(r)*raw = mag_pre - b_mag, and
// mag_pre = e wherever dMag' != 0).  sum_t dMag'[t,f] e[t,f] = sum_c W_dec[c,f] * (DIFF^T E)[c,f]: the unscaled dW_dec
// that the weight-gradient GEMM just produced, so g_r = sum_c W_dec_bf16[c,f] * g_wdec[c,f] - s*b_mag*csum_mag.
// grid ceil(F/32), 256 threads = 8 row lanes x 32 features.
__global__ void gated_rmag_kernel(const float* __restrict__ g_wdec, const bf16* __restrict__ w_dec_bf16,
                                  const float* __restrict__ b_mag, const float* __restrict__ cs_mag, float s, int C,
                                  int F, float* __restrict__ g_r) {
  __shared__ float sh[8][33];
  const int fl = threadIdx.x & 31, cl = threadIdx.x >> 5;
  const int f = blockIdx.x * 32 + fl;
  float acc = 0.f;
  if (f < F)
    for (int c = cl; c < C; c += 8) {
      const size_t i = static_cast<size_t>(c) * F + f;
      acc += __bfloat162float(w_dec_bf16[i]) * g_wdec[i];
    }
  sh[cl][fl] = acc;
  __syncthreads();
  if (cl == 0 && f < F) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][fl];
    g_r[f] = t - s * b_mag[f] * cs_mag[f];
  }
}

int run_prep(cudaStream_t st, const GatedPlan& pl, const svb_gated_params* p, bool train) {
  PrepArgs a{};
  a.w_enc = p->w_gate; a.b_dec = p->b_dec; a.w_enc_bf16 = pl.Wgb; a.dotw = pl.dot;
  a.w_dec = p->w_dec; a.w_dec_bf16 = pl.Wdb;
  a.zero = train ? pl.act_bits : nullptr; a.n_zero = train ? pl.zero_words : 0;
  a.r_mag = p->r_mag; a.exp_r = pl.exp_r;
  a.F = pl.F; a.C = pl.C;
  return run_prep_step(st, a);
}

}  // namespace

extern "C" int svb_gated_forward(svb_handle* h, void* stream, const svb_acts* x, const svb_gated_params* p,
                                 const svb_gated_forward_out* out) {
  if (!h || !out) return fail(SVB_ERR_BAD_ARG, "null handle/out");
  SVB_ON_DEVICE(h);
  SVB_TRY(check_params(x, p));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GatedPlan pl;
  SVB_TRY(plan(h, pl, x, p->F, false));
  h->gradbuf = nullptr;
  const bf16* X = pl.zero_copy_x ? static_cast<const bf16*>(x->x) : pl.X;
  if (!pl.zero_copy_x) SVB_TRY(pack_acts(st, x, pl.X));
  SVB_TRY(run_prep(st, pl, p, false));
  const int T = static_cast<int>(pl.T);
  EpiGatedEnc::Params e1{};
  e1.dot = pl.dot; e1.b_gate = p->b_gate; e1.b_mag = p->b_mag; e1.exp_r = pl.exp_r;
  e1.e_bf16 = (out->enc && out->enc_dtype == SVB_BF16) ? static_cast<bf16*>(out->enc) : pl.E;
  e1.e_f32 = (out->enc && out->enc_dtype == SVB_F32) ? static_cast<float*>(out->enc) : nullptr;
  e1.rp_bf16 = (out->relu_pi && out->relu_pi_dtype == SVB_BF16) ? static_cast<bf16*>(out->relu_pi) : pl.RP;
  e1.rp_f32 = (out->relu_pi && out->relu_pi_dtype == SVB_F32) ? static_cast<float*>(out->relu_pi) : nullptr;
  e1.words = pl.words;
  e1.tma = make_store_tmap_bf16(&e1.tm_e, e1.e_bf16, T, pl.F, pl.F) == 0 &&
           make_store_tmap_bf16(&e1.tm_rp, e1.rp_bf16, T, pl.F, pl.F) == 0;
  SVB_GEMM((launch_gemm_s<false, false, EpiGatedEnc>(st, X, pl.C, pl.Wgb, pl.C, T, pl.F, pl.C, 1, e1)), "gated enc");
  if (out->dec) {
    EpiDec::Params e2{};
    e2.bias = p->b_dec;
    e2.d_bf16 = out->dec_dtype == SVB_BF16 ? static_cast<bf16*>(out->dec) : nullptr;
    e2.d_f32 = out->dec_dtype == SVB_F32 ? static_cast<float*>(out->dec) : nullptr;
    if (e2.d_bf16 && make_store_tmap_bf16(&e2.tm_d, e2.d_bf16, T, pl.C, pl.C)) return fail(SVB_ERR_TMAP, "tensor map");
    SVB_GEMM((launch_gemm_s<false, false, EpiDec>(st, e1.e_bf16, pl.F, pl.Wdb, pl.F, T, pl.C, pl.F, 1, e2)), "dec");
  }
  if (out->via) {
    EpiDec::Params e3{};
    e3.bias = p->b_dec;
    e3.d_bf16 = out->via_dtype == SVB_BF16 ? static_cast<bf16*>(out->via) : nullptr;
    e3.d_f32 = out->via_dtype == SVB_F32 ? static_cast<float*>(out->via) : nullptr;
    if (e3.d_bf16 && make_store_tmap_bf16(&e3.tm_d, e3.d_bf16, T, pl.C, pl.C)) return fail(SVB_ERR_TMAP, "tensor map");
    SVB_GEMM((launch_gemm_s<false, false, EpiDec>(st, e1.rp_bf16, pl.F, pl.Wdb, pl.F, T, pl.C, pl.F, 1, e3)), "via");
  }
  return 0;
}

extern "C" int svb_gated_step_grads(svb_handle* h, void* stream, const svb_acts* x, const svb_gated_params* p,
                                    float lambda_sparse, int64_t global_tokens, const svb_train_out* out) {
  if (!h) return fail(SVB_ERR_BAD_ARG, "null handle");
  SVB_ON_DEVICE(h);
  SVB_TRY(check_params(x, p));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GatedPlan pl;
  SVB_TRY(plan(h, pl, x, p->F, true));
  const int T = static_cast<int>(pl.T), C = pl.C, F = pl.F;
  const double Tg = global_tokens > 0 ? static_cast<double>(global_tokens) : static_cast<double>(pl.T);
  const bf16* X = pl.zero_copy_x ? static_cast<const bf16*>(x->x) : pl.X;
  h->step_flags = 0;
  prof_begin_step(h);   // phases as in svb_sae.cu; "dec_gemm" covers the decoder and the via_gate GEMM
  prof_mark(h, st, 0);
  // slab-major X / DIFF + fused NCHW decoder epilogue under the same conditions as svb_sae_step_grads
  void* dec_out = out ? out->dec_out : nullptr;
  const bool slab_ok = !pl.zero_copy_x && post_dec_fusable(x, dec_out, out ? out->dec_layout : SVB_NCHW);
  // zero-copy token-major (channels_last) activations through the fused decoder epilogue, as in svb_sae_step_grads
  const bool tok_fused = pl.zero_copy_x && x->layout == SVB_TOKENS && pl.hw >= 32 && x->n_images <= 65535 &&
                         (!dec_out || (out->dec_layout == SVB_TOKENS && out->dec_dtype == SVB_BF16 &&
                                       (reinterpret_cast<uintptr_t>(dec_out) & 15) == 0));
  pl.fused_dec = (slab_ok && pl.hw >= 32) || tok_fused;
  pl.xs = slab_ok && (C % 64 == 0 || pl.fused_dec);
  const bool ds = pl.xs || tok_fused;   // DIFF slab-major
  int out_kind = 0;
  if (dec_out && tok_fused) out_kind = 2;
  else if (dec_out)
    out_kind = (out->dec_dtype == SVB_BF16 && pl.hw % 8 == 0 && (reinterpret_cast<uintptr_t>(dec_out) & 15) == 0) ? 1 : 4;
  const long long ld_t = (pl.T + 7) & ~7LL;
  // weight prologue on the side stream, next to the activation pack (svb_common.cuh: side_fork / side_join)
  SVB_TRY(side_fork(h, st));
  SVB_TRY(run_prep(h->side, pl, p, true));
  if (!pl.zero_copy_x) SVB_TRY(pack_acts(st, x, pl.X, pl.xs, pl.fused_dec ? pl.xpart : nullptr));
  if (tok_fused) {  // zero-copy tokens: only their statistics are needed (svb_sae.cu)
    launch_x_stats_tokens(st, X, pl.xpart, C, pl.hw, pl.nt_hw, pl.n_img);
    SVB_LAUNCH_CHECK("x_stats_tokens");
  }
  SVB_TRY(side_join(h, st));

  prof_mark(h, st, 1);
  EpiGatedEnc::Params e1{};
  e1.dot = pl.dot; e1.b_gate = p->b_gate; e1.b_mag = p->b_mag; e1.exp_r = pl.exp_r;
  e1.e_bf16 = pl.E; e1.rp_bf16 = pl.RP; e1.l1_partial = pl.l1_part;
  e1.mask_e = pl.mask_e; e1.mask_rp = pl.mask_rp;
  e1.words = pl.words; e1.tma = 1; e1.slab_major = pl.es;
  if (pl.es ? (make_store_tmap_bf16_slab(&e1.tm_e, pl.E, T, F) || make_store_tmap_bf16_slab(&e1.tm_rp, pl.RP, T, F))
            : (make_store_tmap_bf16(&e1.tm_e, pl.E, T, F, F) || make_store_tmap_bf16(&e1.tm_rp, pl.RP, T, F, F)))
    return fail(SVB_ERR_TMAP, "tensor maps for E / relu_pi");
  SVB_GEMM((launch_gemm_s<false, false, EpiGatedEnc>(st, X, C, pl.Wgb, C, T, F, C, 1, e1, nullptr, 0, 0, pl.xs, false)), "gated enc");
  prof_mark(h, st, 2);
  // per-image activity bits of e from its 1-bit mask: side stream, beside the decoder GEMM
  SVB_TRY(side_fork(h, st));
  (mask_to_activity_kernel<<<dim3(static_cast<unsigned>(pl.n_img), cdiv(pl.words, 4)), 128, 0, h->side>>>(
      pl.mask_e, pl.act_bits, pl.T, pl.hw, pl.words), svb::count_launch());
  EpiDec::Params e2v{};
  e2v.bias = p->b_dec; e2v.x = X; e2v.sq_partial = pl.aux_part; e2v.x_slab = pl.xs;   // via_gate: aux loss value only
  if (pl.fused_dec) {
    // decoder epilogue writes NCHW d, DIFF (slab-major) and the per-image channel statistics itself (EpiDecNchw)
    EpiDecNchw::Params e2{};
    e2.bias = p->b_dec; e2.x = X; e2.sq_partial = pl.sq_part; e2.part = pl.dpart; e2.hw = pl.hw;
    e2.out = dec_out; e2.out_kind = out_kind; e2.x_slab = pl.xs ? 1 : 0; e2.tok = tok_fused ? 1 : 0;
    if (make_store_tmap_bf16_slab32(&e2.tm_diff, pl.DIFF, T, C)) return fail(SVB_ERR_TMAP, "tensor map for DIFF");
    if (out_kind == 2 && make_store_tmap_bf16_chunk(&e2.tm_out, dec_out, T, C, C)) return fail(SVB_ERR_TMAP, "tensor map for the token-major output");
    if (out_kind == 1 && make_tmap_nchw_bf16(&e2.tm_out, dec_out, pl.n_img, C, pl.hw)) return fail(SVB_ERR_TMAP, "tensor map for the NCHW output");
    if (out_kind == 4 && make_store_tmap_bf16_cmajor(&e2.tm_out, pl.D, C, pl.T, ld_t)) return fail(SVB_ERR_TMAP, "tensor map for the channel-major output");
    // newest E tiles first (still in L2), as in svb_sae.cu
    SVB_GEMM((launch_gemm_s<false, false, EpiDecNchw>(st, pl.E, F, pl.Wdb, F, T, C, F, 1, e2, nullptr, 0, 0, pl.es, false, 0, /*reverse_m=*/true)), "dec (fused NCHW)");
    // the statistics folds (and the scatter of a channel-major d) only feed the end of the step: side stream
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
    // statistics + NCHW write-back of d only feed the end of the step: side stream, beside the via / dE GEMMs
    SVB_TRY(side_fork(h, st));
    SVB_TRY(run_post_dec(h->side, x, X, pl.D, pl.T, dec_out, out ? out->dec_dtype : SVB_BF16,
                         out ? out->dec_layout : SVB_NCHW, pl.st, pl.chan, pl.var_part, pl.rowvar, pl.xs));
  }
  SVB_GEMM((launch_gemm_s<false, false, EpiDec>(st, pl.RP, F, pl.Wdb, F, T, C, F, 1, e2v, nullptr, 0, 0, pl.es, false)), "via");
  prof_mark(h, st, 3);
  prof_mark(h, st, 4);
  EpiGatedDPre::Params e3{};
  e3.mask_e = pl.mask_e; e3.mask_rp = pl.mask_rp; e3.exp_r = pl.exp_r; e3.words = pl.words;
  e3.colsum_mag = pl.cs_mag; e3.colsum_a = pl.cs_a;
  e3.l1c = static_cast<float>(static_cast<double>(lambda_sparse) * C / (2.0 * F));
  e3.block_n = 256; e3.slab_major = pl.es;
  if (pl.es ? make_store_tmap_bf16_slab(&e3.tm_a, pl.A, T, F) : make_store_tmap_bf16(&e3.tm_a, pl.A, T, F, F))
    return fail(SVB_ERR_TMAP, "tensor map for A'");
  // (single-CTA: on SM pairs this epilogue-heavy GEMM measured 0.424 against 0.398 ms at cfg3)
  SVB_GEMM((launch_gemm<256, false, true, EpiGatedDPre>(st, pl.DIFF, C, pl.Wdb, F, T, F, C, 1, e3, nullptr, 0, 0, ds, false)), "gated dE");
  prof_mark(h, st, 5);
  // Weight gradients: the gate side first, so that [gW_gate | gb_gate | gb_mag | gr_mag] can be all-reduced while the
  // decoder weight-gradient GEMM runs (svb_set_comm_stream).
  const size_t FC = static_cast<size_t>(F) * C;
  const float s = static_cast<float>(2.0 / (Tg * C));
  float* flat = pl.flat;
  EpiPartial::Params e5{pl.P_wg, C, static_cast<long long>(FC)};
  SVB_GEMM((launch_gemm_s<true, true, EpiPartial>(st, pl.A, F, X, C, F, C, T, 0, e5, nullptr, 0, 0, pl.es, pl.xs)), "dW_gate");
  // reductions + gate-side assembly + tail on the side stream, beside the dW_dec GEMM
  SVB_TRY(side_fork(h, st));
  cudaStream_t ss = h->side;
  SVB_TRY(reduce_rows(ss, pl.cs_mag, pl.tiles_m, F, 1.f, pl.stage, pl.csum_mag));
  SVB_TRY(reduce_rows(ss, pl.cs_a, pl.tiles_m, F, 1.f, pl.stage, pl.csum_a));
  (mask_colcount_kernel<<<dim3(cdiv(pl.words, 128), pl.cnt_chunks), 128, 0, ss>>>(pl.mask_rp, pl.T, pl.words, pl.cnt_part), svb::count_launch());
  (gated_vec_grads_kernel<<<cdiv(F, 256), 256, 0, ss>>>(pl.csum_mag, pl.cnt_part, pl.cnt_chunks, pl.words * 32, e3.l1c, s, F,
                                                      flat + pl.o_gbg, flat + pl.o_gbm), svb::count_launch());
  AssembleArgs aa{};
  aa.P_wd = pl.P_wd; aa.g_wdec = flat + pl.o_gwd; aa.s_wd = pl.s_wd;
  aa.P_we = pl.P_wg; aa.g_wenc = flat + pl.o_gwg; aa.s_we = pl.s_wg;
  aa.csum = pl.csum_a; aa.b_dec = p->b_dec; aa.g_benc = nullptr;  // the three vector gradients are written above
  aa.w_enc_bf16 = pl.Wgb; aa.vm = pl.vm; aa.vm_chunks = kVmChunks;
  aa.act_bits = pl.act_bits; aa.count = flat + pl.o_count; aa.n_active = out ? out->activity.n_active : nullptr;
  aa.nact_f = pl.nact_f; aa.n_img = static_cast<int>(pl.n_img); aa.words = pl.words;
  aa.F = F; aa.C = C; aa.s = s;
  SVB_TRY(run_assemble(ss, aa, 1));
  SVB_TRY(release_comm_stream(h, ss));
  h->early_elems = h->comm ? static_cast<int64_t>(pl.o_gr) : 0;   // gr_mag needs the decoder weight gradient
  TailArgs ta{};
  ta.chan = pl.chan; ta.vm = pl.vm; ta.vm_chunks = kVmChunks; ta.g_bdec = flat + pl.o_gbd; ta.s = s;
  ta.sq_part = pl.sq_part; ta.n_sq = pl.sms * 8;
  ta.l1_part = pl.l1_part; ta.n_l1 = pl.sms * 8;
  ta.aux_part = pl.aux_part; ta.n_aux = pl.sms * 8;
  ta.nact_f = pl.nact_f; ta.n_img = static_cast<int>(pl.n_img);
  ta.var_part = pl.var_part; ta.n_var_part = pl.fused_dec ? cdiv(C, 32) : cdiv(C, 8); ta.rowvar = pl.rowvar; ta.n_rows = pl.hw == 1 ? pl.T : 0;
  ta.flat = flat; ta.o_sums = pl.o_sums; ta.o_chansq = pl.o_chansq; ta.o_max = pl.o_max; ta.C = C;
  (grads_tail_kernel<<<1, 1024, 0, ss>>>(ta), svb::count_launch());
  prof_mark(h, st, 6);
  EpiPartial::Params e4{pl.P_wd, F, static_cast<long long>(FC)};
  SVB_GEMM((launch_gemm_s<true, true, EpiPartial>(st, pl.DIFF, C, pl.E, F, C, F, T, 0, e4, nullptr, 0, 0, ds, pl.es)), "dW_dec");
  prof_mark(h, st, 7);
  SVB_TRY(side_join(h, st));
  SVB_TRY(run_assemble(st, aa, 2));
  (gated_rmag_kernel<<<cdiv(F, 32), 256, 0, st>>>(flat + pl.o_gwd, pl.Wdb, p->b_mag, pl.csum_mag, s, C, F, flat + pl.o_gr), svb::count_launch());
  SVB_LAUNCH_CHECK("gated grad assembly");
  prof_mark(h, st, 8);
  h->gradbuf = flat;
  h->sum_elems = static_cast<int64_t>(pl.sum_elems);
  h->max_elems = static_cast<int64_t>(pl.max_elems);
  return 0;
}

extern "C" int svb_gated_step_apply(svb_handle* h, void* stream, const svb_acts* x, const svb_gated_params* p,
                                    const svb_adam_state* adam, const svb_opt_config* opt, float lambda_sparse,
                                    int32_t expansion_factor, int64_t global_tokens, int64_t global_images,
                                    const svb_train_out* out) {
  if (!h || !adam || !opt) return fail(SVB_ERR_BAD_ARG, "null handle/adam/opt");
  SVB_ON_DEVICE(h);
  SVB_TRY(check_params(x, p));
  for (int i = 0; i < 6; ++i)
    if (!adam->m[i] || !adam->v[i]) return fail(SVB_ERR_BAD_ARG, "null Adam state tensor %d", i);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GatedPlan pl;
  SVB_TRY(plan(h, pl, x, p->F, true));
  if (pl.flat != h->gradbuf) return fail(SVB_ERR_BAD_ARG, "svb_gated_step_apply called without a matching svb_gated_step_grads");
  const int C = pl.C, F = pl.F;
  const size_t FC = static_cast<size_t>(F) * C;
  float* flat = pl.flat;
  AdamCoef k;
  SVB_TRY(adam_coef_for(h, st, opt, &k));
  AdamSeg segs[6];
  int ns = 0;
  segs[ns++] = AdamSeg{p->w_gate, flat + pl.o_gwg, adam->m[0], adam->v[0], FC};
  segs[ns++] = AdamSeg{p->b_gate, flat + pl.o_gbg, adam->m[1], adam->v[1], static_cast<unsigned long long>(F)};
  segs[ns++] = AdamSeg{p->b_mag, flat + pl.o_gbm, adam->m[2], adam->v[2], static_cast<unsigned long long>(F)};
  segs[ns++] = AdamSeg{p->r_mag, flat + pl.o_gr, adam->m[3], adam->v[3], static_cast<unsigned long long>(F)};
  segs[ns++] = AdamSeg{p->b_dec, flat + pl.o_gbd, adam->m[5], adam->v[5], static_cast<unsigned long long>(C)};
  if (opt->optimizer != SVB_CONSTRAINED_ADAM) segs[ns++] = AdamSeg{p->w_dec, flat + pl.o_gwd, adam->m[4], adam->v[4], FC};
  // independent pieces of the tail side by side, as in svb_sae_step_apply
  const bool cadam = opt->optimizer == SVB_CONSTRAINED_ADAM;
  SVB_TRY(side_fork(h, st));
  cudaStream_t s2 = h->side;
  SVB_TRY(run_adam_multi(cadam ? s2 : st, segs, ns, k));
  if (cadam) launch_cadam(st, p->w_dec, flat + pl.o_gwd, adam->m[4], adam->v[4], C, F, k);
  SVB_LAUNCH_CHECK("gated adam");
  if (out && (out->stats || out->activity.dead || out->activity.freq)) {
    FinalizeArgs fa{};
    fa.flat = flat; fa.o_sums = pl.o_sums; fa.o_chansq = pl.o_chansq; fa.o_max = pl.o_max; fa.o_count = pl.o_count;
    fa.C = C; fa.F = F; fa.expansion = expansion_factor;
    fa.T_g = static_cast<float>(global_tokens > 0 ? global_tokens : pl.T);
    fa.B_g = static_cast<float>(global_images > 0 ? global_images : pl.n_img);
    fa.lambda = lambda_sparse;   // utils.py:2473: loss = rec + lambda * l1 + aux
    fa.stats = out->stats; fa.dead = out->activity.dead; fa.freq = out->activity.freq;
    (step_finalize_kernel<<<1, 1024, 0, s2>>>(fa), svb::count_launch());
    SVB_LAUNCH_CHECK("gated finalize");
  }
  SVB_TRY(side_join(h, st));
  prof_mark(h, st, 9);
  return 0;
}

extern "C" int svb_gated_train_step(svb_handle* h, void* stream, const svb_acts* x, const svb_gated_params* p,
                                    const svb_adam_state* adam, const svb_opt_config* opt, float lambda_sparse,
                                    int32_t expansion_factor, const svb_train_out* out) {
  SVB_TRY(svb_gated_step_grads(h, stream, x, p, lambda_sparse, 0, out));
  return svb_gated_step_apply(h, stream, x, p, adam, opt, lambda_sparse, expansion_factor, 0, 0, out);
}
