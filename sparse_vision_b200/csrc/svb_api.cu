// Handle lifecycle, optimiser / re-initialisation / activity entry points, layout helpers and the indirect-effect
// reductions of the C ABI (include/svb.h).
#include "svb_common.cuh"
#include "gemm2_sm100.cuh"
#include "fused_ie_sm100.cuh"

using namespace svb;

extern "C" const char* svb_last_error(void) { return err_buf(); }
extern "C" int svb_version(void) { return 100; }

extern "C" int svb_create(svb_handle** out) {
  if (!out) return fail(SVB_ERR_BAD_ARG, "null out pointer");
  int dev = 0;
  SVB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  SVB_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(SVB_ERR_UNSUPPORTED, "device %d is sm_%d%d; libsvb is built for sm_100a (B200) only", dev, prop.major,
                prop.minor);
  if (!tmap_encode_fn()) return fail(SVB_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  svb_handle* h = new svb_handle();
  h->device = dev;
  h->sms = prop.multiProcessorCount;
  if (cudaMalloc(&h->coef_dev, 2 * sizeof(float)) != cudaSuccess) {   // here, not lazily: never inside a stream capture
    cudaGetLastError();
    delete h;
    return fail(SVB_ERR_NOMEM, "cudaMalloc failed in svb_create");
  }
  *out = h;
  return 0;
}

extern "C" int svb_destroy(svb_handle* h) {
  if (!h) return 0;
  if (h->arena.base) cudaFree(h->arena.base);
  if (h->ev_early) cudaEventDestroy(h->ev_early);
  if (h->side) { cudaStreamDestroy(h->side); cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join); }
  if (h->coef_dev) cudaFree(h->coef_dev);
  svb_comm_destroy(h);
  delete h;
  return 0;
}

extern "C" int64_t svb_launch_count(void) { return static_cast<int64_t>(launch_counter()); }
extern "C" int32_t svb_last_step_flags(const svb_handle* h) { return h ? h->step_flags : 0; }
extern "C" int svb_set_tuning(int32_t key, int32_t value) {
  if (key < 0 || key >= kTuneCount) return fail(SVB_ERR_BAD_ARG, "unknown tuning key %d", key);
  tuning(key) = value;
  return 0;
}
extern "C" int32_t svb_get_tuning(int32_t key) { return (key < 0 || key >= kTuneCount) ? -1 : tuning(key); }

// ---------------------------------------------------------------------------------------------------- profiling
static const char* kPhaseNames[] = {"pack+prep", "enc_gemm", "dec_gemm", "channel_stats", "dE_gemm",
                                    "dWenc_gemm", "dWdec_gemm", "grad_assembly", "adam+finalize"};

extern "C" int svb_set_comm_stream(svb_handle* h, void* stream) {
  if (!h) return fail(SVB_ERR_BAD_ARG, "null handle");
  h->comm = static_cast<cudaStream_t>(stream);
  return 0;
}
extern "C" int svb_grad_early_elems(svb_handle* h, int64_t* elems) {
  if (!h || !elems) return fail(SVB_ERR_BAD_ARG, "null argument");
  *elems = h->gradbuf ? h->early_elems : 0;
  return 0;
}

extern "C" int svb_profile_enable(svb_handle* h, int32_t enable) {
  if (!h) return fail(SVB_ERR_BAD_ARG, "null handle");
  SVB_ON_DEVICE(h);
  Profiler& p = h->prof;
  if (enable && !p.created) {
    for (int s = 0; s < kProfMaxSteps; ++s)
      for (int m = 0; m < kProfMaxMarks; ++m) SVB_CUDA(cudaEventCreate(&p.ev[s][m]));
    p.created = true;
  }
  p.on = enable != 0;
  p.step = -1;
  p.steps_recorded = 0;
  return 0;
}

extern "C" int svb_profile_read(svb_handle* h, int32_t max_phases, float* ms_avg_host, int32_t* n_phases,
                                int32_t* n_steps) {
  if (!h || !ms_avg_host || !n_phases || !n_steps) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  Profiler& p = h->prof;
  if (!p.created) return fail(SVB_ERR_BAD_ARG, "profiling was never enabled");
  SVB_CUDA(cudaDeviceSynchronize());
  const int ns = static_cast<int>(p.steps_recorded < kProfMaxSteps ? p.steps_recorded : kProfMaxSteps);
  int phases = 0;
  double acc[kProfMaxMarks] = {0};
  int used = 0;
  for (int s = 0; s < ns; ++s) {
    const int marks = p.marks[s];
    if (marks < 2) continue;
    if (marks - 1 > phases) phases = marks - 1;
    for (int m = 0; m + 1 < marks; ++m) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, p.ev[s][m], p.ev[s][m + 1]) == cudaSuccess) acc[m] += ms;
    }
    ++used;
  }
  if (phases > max_phases) phases = max_phases;
  for (int m = 0; m < phases; ++m) ms_avg_host[m] = used ? static_cast<float>(acc[m] / used) : 0.f;
  *n_phases = phases;
  *n_steps = used;
  return 0;
}

extern "C" const char* svb_profile_phase_name(int32_t i) {
  return (i >= 0 && i < static_cast<int>(sizeof(kPhaseNames) / sizeof(kPhaseNames[0]))) ? kPhaseNames[i] : "";
}

extern "C" int64_t svb_workspace_bytes(const svb_handle* h) { return h ? static_cast<int64_t>(h->arena.cap) : 0; }

// ---------------------------------------------------------------------------------------------------- optimiser
extern "C" int svb_adam_step(svb_handle* h, void* stream, int32_t n_tensors, float* const* params,
                             const float* const* grads, float* const* m, float* const* v, const int64_t* rows,
                             const int64_t* cols, int32_t decoder_index, const svb_opt_config* opt) {
  if (!h || !params || !grads || !m || !v || !rows || !cols || !opt) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AdamCoef k;
  SVB_TRY(adam_coef_for(h, st, opt, &k));
  for (int i = 0; i < n_tensors; ++i) {
    if (!params[i] || !m[i] || !v[i]) return fail(SVB_ERR_BAD_ARG, "null tensor %d", i);
    if (!grads[i]) continue;  // parameter without gradient: torch.optim.Adam skips it
    const size_t n = static_cast<size_t>(rows[i]) * cols[i];
    if (i == decoder_index && opt->optimizer == SVB_CONSTRAINED_ADAM) {
      launch_cadam(st, params[i], const_cast<float*>(grads[i]), m[i], v[i], static_cast<int>(rows[i]),
                   static_cast<int>(cols[i]), k);
    } else {
      (adam_kernel<<<grid_for(n), 256, 0, st>>>(params[i], grads[i], m[i], v[i], n, k, nullptr), svb::count_launch());
    }
  }
  // ConstrainedAdam renormalises the decoder columns even when it had no gradient (utils.py:76-79)
  if (decoder_index >= 0 && decoder_index < n_tensors && opt->optimizer == SVB_CONSTRAINED_ADAM &&
      !grads[decoder_index])
    (renorm_columns_kernel<<<cdiv(cols[decoder_index], 32), 256, 0, st>>>(
        params[decoder_index], static_cast<int>(rows[decoder_index]), static_cast<int>(cols[decoder_index])), svb::count_launch());
  SVB_LAUNCH_CHECK("adam_step");
  return 0;
}

extern "C" int svb_reinit_dead(svb_handle* h, void* stream, const svb_sae_params* p, int32_t C,
                               const svb_adam_state* adam, const uint8_t* dead, const float* new_w_enc,
                               const float* new_w_dec, float new_b_enc) {
  if (!h || !p || !dead || !new_w_enc || !new_w_dec) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = static_cast<size_t>(p->F) * C;
  (reinit_scatter_kernel<<<grid_for(n), 256, 0, st>>>(dead, p->F, C, p->w_enc, p->b_enc, p->w_dec, new_w_enc, new_w_dec,
                                                     new_b_enc, adam ? adam->m[0] : nullptr, adam ? adam->v[0] : nullptr,
                                                     adam ? adam->m[1] : nullptr, adam ? adam->v[1] : nullptr,
                                                     adam ? adam->m[2] : nullptr, adam ? adam->v[2] : nullptr), svb::count_launch());
  (renorm_columns_kernel<<<cdiv(p->F, 32), 256, 0, st>>>(p->w_dec, C, p->F), svb::count_launch());
  SVB_LAUNCH_CHECK("reinit_dead");
  return 0;
}

// ---------------------------------------------------------------------------------------------------- activity
extern "C" int svb_measure_inactive(svb_handle* h, void* stream, const void* t, int32_t dtype, int32_t layout,
                                    int64_t n_images, int32_t hw, int32_t F, const svb_activity_out* act) {
  if (!h || !t || !act) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (n_images <= 0 || hw <= 0 || F <= 0) return fail(SVB_ERR_BAD_ARG, "empty tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int words = (F + 31) / 32;
  // rows of the bit matrix: images for NCHW, tokens for a 2-D tensor
  const long long n_rows = (layout == SVB_NCHW) ? n_images : n_images * hw;
  Arena dry; dry.dry = true;
  dry.take<uint32_t>(static_cast<size_t>(n_rows) * words); dry.take<float>(F); dry.take<float>(n_rows);
  SVB_TRY(ensure_arena(h, dry.off));
  h->arena.off = 0; h->arena.dry = false;
  uint32_t* bits = h->arena.take<uint32_t>(static_cast<size_t>(n_rows) * words);
  float* count = h->arena.take<float>(F);
  float* nact = h->arena.take<float>(n_rows);
  h->gradbuf = nullptr;
  if (n_rows >= (1LL << 31)) return fail(SVB_ERR_UNSUPPORTED, "too many rows");
  if (!(layout == SVB_NCHW && hw > 1) && dtype == SVB_BF16 && F % 8 == 0 && n_rows >= 1024 && n_rows < (1LL << 24) &&
      (reinterpret_cast<uintptr_t>(t) & 15) == 0 && cdiv(n_rows, kActRows) <= 65535) {
    // many rows (the 2-D call on a token-major encoder output: every token is a sample): one fused pass, no bit matrix
    SVB_CUDA(cudaMemsetAsync(count, 0, sizeof(float) * F, st));
    if (act->n_active) SVB_CUDA(cudaMemsetAsync(act->n_active, 0, sizeof(int32_t) * n_rows, st));
    (activity_rows_fused_kernel<<<dim3(cdiv(F / 8, 256), cdiv(n_rows, kActRows)), 256, 0, st>>>(
         static_cast<const uint4*>(t), n_rows, F / 8, count, act->n_active), svb::count_launch());
    (activity_finalize_kernel<<<1, 1024, 0, st>>>(count, F, static_cast<float>(n_rows), act->dead, act->freq, nullptr), svb::count_launch());
    SVB_LAUNCH_CHECK("measure_inactive (fused rows)");
    return 0;
  }
  if (layout == SVB_NCHW && hw > 1) {
    (fill_u32_kernel<<<grid_for(static_cast<size_t>(n_rows) * words), 256, 0, st>>>(bits, static_cast<size_t>(n_rows) * words, 0u), svb::count_launch());
    const long long warps = n_images * F;
    if (dtype == SVB_F32)
      (activity_bits_nchw_kernel<float><<<cdiv(warps, 8), 256, 0, st>>>(static_cast<const float*>(t), bits, static_cast<int>(n_images), F, hw, words), svb::count_launch());
    else
      (activity_bits_nchw_kernel<bf16><<<cdiv(warps, 8), 256, 0, st>>>(static_cast<const bf16*>(t), bits, static_cast<int>(n_images), F, hw, words), svb::count_launch());
  } else {
    const long long n = n_rows * words;
    if (dtype == SVB_F32)
      (activity_bits_rows_kernel<float><<<cdiv(n, 256), 256, 0, st>>>(static_cast<const float*>(t), bits, n_rows, F, words), svb::count_launch());
    else
      (activity_bits_rows_kernel<bf16><<<cdiv(n, 256), 256, 0, st>>>(static_cast<const bf16*>(t), bits, n_rows, F, words), svb::count_launch());
  }
  (activity_count_kernel<<<words, 256, 0, st>>>(bits, static_cast<int>(n_rows), words, F, count), svb::count_launch());
  (activity_per_image_kernel<<<cdiv(n_rows, 8), 256, 0, st>>>(bits, static_cast<int>(n_rows), words, act->n_active, nact), svb::count_launch());
  (activity_finalize_kernel<<<1, 1024, 0, st>>>(count, F, static_cast<float>(n_rows), act->dead, act->freq, nullptr), svb::count_launch());
  SVB_LAUNCH_CHECK("measure_inactive");
  return 0;
}

// ---------------------------------------------------------------------------------------------------- layout
extern "C" int svb_pack_tokens(svb_handle* h, void* stream, const svb_acts* x, void* out_bf16_tokens) {
  if (!h || !out_bf16_tokens) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  SVB_TRY(check_acts(x));
  return pack_acts(static_cast<cudaStream_t>(stream), x, static_cast<bf16*>(out_bf16_tokens));
}

extern "C" int svb_unpack_tokens(svb_handle* h, void* stream, const void* tokens, int32_t tokens_dtype,
                                 int64_t n_images, int32_t hw, int32_t C, void* out_nchw, int32_t out_dtype) {
  if (!h || !tokens || !out_nchw) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (tokens_dtype != SVB_BF16) return fail(SVB_ERR_UNSUPPORTED, "svb_unpack_tokens takes bf16 tokens");
  return unpack_to(static_cast<cudaStream_t>(stream), static_cast<const bf16*>(tokens), n_images, hw, C, out_nchw,
                   out_dtype, SVB_NCHW);
}

// ---------------------------------------------------------------------------------------------------- IE
namespace {

int ie_chunks(long long T, int col_tiles, int sms) {
  long long want = (4LL * sms + col_tiles - 1) / col_tiles;
  long long max_chunks = (T + 63) / 64;  // at least 64 rows per chunk (8 warps x 8 rows)
  if (want > max_chunks) want = max_chunks;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  return static_cast<int>(want);
}

// a, g token-major [T,F]; avgT [HW,F]; out[f] = scale * sum.   `partial` holds chunks*F floats, `stage` 32*F.
template <typename T>
int launch_ie_channelwise(cudaStream_t st, int sms, const T* a, const T* g, const float* avgT, long long Tn, int HW,
                          int F, float scale, float* partial, int chunks, float* stage, float* out) {
  constexpr int V = Vec16<T>::kN;
  dim3 grid(cdiv(F, 32 * V), chunks);
  (ie_channelwise_kernel<T, 4><<<grid, 256, 0, st>>>(a, g, avgT, Tn, HW, F, partial), svb::count_launch());
  SVB_LAUNCH_CHECK("ie_channelwise");
  return reduce_rows(st, partial, chunks, F, scale, stage, out);
}

}  // namespace

extern "C" int svb_ie_channelwise(svb_handle* h, void* stream, const void* a, const void* g, int32_t dtype,
                                  const float* avg, int64_t n_images, int32_t hw, int32_t F, float scale, float* out) {
  if (!h || !a || !g || !avg || !out) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (n_images <= 0 || hw <= 0 || F <= 0) return fail(SVB_ERR_BAD_ARG, "empty input");
  const int V = dtype == SVB_F32 ? 4 : 8;
  if (F % V) return fail(SVB_ERR_UNSUPPORTED, "F=%d must be a multiple of %d for 16-byte loads", F, V);
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(g)) & 15)
    return fail(SVB_ERR_BAD_ARG, "a / g must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long T = n_images * static_cast<long long>(hw);
  const int chunks = ie_chunks(T, cdiv(F, 32 * V), h->sms);
  Arena dry; dry.dry = true;
  dry.take<float>(static_cast<size_t>(hw) * F); dry.take<float>(static_cast<size_t>(chunks) * F); dry.take<float>(32 * static_cast<size_t>(F));
  SVB_TRY(ensure_arena(h, dry.off));
  h->arena.off = 0; h->arena.dry = false; h->gradbuf = nullptr;
  float* avgT = h->arena.take<float>(static_cast<size_t>(hw) * F);
  float* partial = h->arena.take<float>(static_cast<size_t>(chunks) * F);
  float* stage = h->arena.take<float>(32 * static_cast<size_t>(F));
  (transpose_f32_kernel<<<dim3(cdiv(hw, 32), cdiv(F, 32)), dim3(32, 8), 0, st>>>(avg, avgT, F, hw), svb::count_launch());
  SVB_LAUNCH_CHECK("transpose avg");
  if (dtype == SVB_F32)
    return launch_ie_channelwise<float>(st, h->sms, static_cast<const float*>(a), static_cast<const float*>(g), avgT, T,
                                        hw, F, scale, partial, chunks, stage, out);
  if (dtype == SVB_BF16)
    return launch_ie_channelwise<bf16>(st, h->sms, static_cast<const bf16*>(a), static_cast<const bf16*>(g), avgT, T, hw,
                                       F, scale, partial, chunks, stage, out);
  return fail(SVB_ERR_BAD_ARG, "bad dtype %d", dtype);
}

extern "C" int svb_ie_allchannels(svb_handle* h, void* stream, const void* err, const void* g, int32_t dtype,
                                  const float* avg, int64_t n_images, int32_t C, int32_t hw, float scale, float* out) {
  if (!h || !err || !g || !avg || !out) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (n_images <= 0 || hw <= 0 || C <= 0) return fail(SVB_ERR_BAD_ARG, "empty input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n_pix = n_images * static_cast<long long>(hw);
  const int blocks = cdiv(n_pix, 32);
  Arena dry; dry.dry = true;
  dry.take<float>(blocks);
  SVB_TRY(ensure_arena(h, dry.off));
  h->arena.off = 0; h->arena.dry = false; h->gradbuf = nullptr;
  float* partial = h->arena.take<float>(blocks);
  if (dtype == SVB_F32)
    (ie_allchannels_nchw_kernel<float><<<blocks, dim3(32, 8), 0, st>>>(static_cast<const float*>(err), static_cast<const float*>(g), avg, n_pix, C, hw, partial), svb::count_launch());
  else if (dtype == SVB_BF16)
    (ie_allchannels_nchw_kernel<bf16><<<blocks, dim3(32, 8), 0, st>>>(static_cast<const bf16*>(err), static_cast<const bf16*>(g), avg, n_pix, C, hw, partial), svb::count_launch());
  else
    return fail(SVB_ERR_BAD_ARG, "bad dtype %d", dtype);
  (reduce_flat_kernel<<<1, 1024, 0, st>>>(partial, static_cast<size_t>(blocks), scale, out), svb::count_launch());
  SVB_LAUNCH_CHECK("ie_allchannels");
  return 0;
}

// Node-IE for one layer.  C % 128 == 0, C <= 256: ONE fused kernel keeps a and G = g W_dec in TMEM and reduces them on the
// spot (fused_ie_sm100.cuh), plus the two small [T, C] passes.  Other shapes: encoder GEMM, decoder GEMM, G = g W_dec GEMM,
// then the three reductions on bf16 tokens.
extern "C" int svb_node_ie_layer(svb_handle* h, void* stream, const svb_acts* x, const void* grad,
                                 const svb_sae_params* p, const float* enc_avg, const float* err_avg,
                                 const float* x_avg, float scale, float* ie_features, float* ie_error,
                                 float* ie_neurons) {
  if (!h || !grad || !p || !enc_avg || !err_avg || !x_avg) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  SVB_TRY(check_acts(x));
  if (!p->w_enc || !p->b_enc || !p->w_dec || !p->b_dec || p->F <= 0 || p->F % 8)
    return fail(SVB_ERR_BAD_ARG, "bad SAE parameters");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long T = x->n_images * static_cast<long long>(x->hw);
  const int C = x->C, F = p->F, HW = x->hw, Ti = static_cast<int>(T);
  const size_t TC = static_cast<size_t>(T) * C, TF = static_cast<size_t>(T) * F, FC = static_cast<size_t>(F) * C;
  const int chunks_f = ie_chunks(T, cdiv(F, 256), h->sms), chunks_c = ie_chunks(T, cdiv(C, 256), h->sms);
  const int tok_blocks = cdiv(T, 8);
  svb_acts gx = *x;
  gx.x = grad;
  const bool zx = acts_are_bf16_tokens(x), zg = acts_are_bf16_tokens(&gx);
  // Fused path (fused_ie_sm100.cuh): a and G stay in TMEM, no decoder GEMM; C % 128 == 0, C <= 256
  const bool fused = tuning(kTuneFusedIe) != 0 && fused_ie_supported(T, C, F, h->sms);
  const int ie_slots = fused ? fused_ie_slots(T, F, h->sms) : 0, q_rows = fused ? fused_ie_qrows(F) : 0;
  bf16 *X = nullptr, *G = nullptr, *Web, *Wdb, *E = nullptr, *GE = nullptr, *DIFF = nullptr;
  float *fold, *avgT_f = nullptr, *avgT_c, *avgT_e, *partial, *stage, *tokpart, *ie_part = nullptr, *q_part = nullptr;
  auto carve = [&](Arena& ar) {
    if (!zx) X = ar.take<bf16>(TC);
    if (!zg) G = ar.take<bf16>(TC);
    Web = ar.take<bf16>(FC); Wdb = ar.take<bf16>(FC); fold = ar.take<float>(F);
    if (fused) {
      ie_part = ar.take<float>(static_cast<size_t>(2 * ie_slots) * F);
      q_part = ar.take<float>(static_cast<size_t>(q_rows) * T);
    } else {
      E = ar.take<bf16>(TF); GE = ar.take<bf16>(TF); DIFF = ar.take<bf16>(TC);
      avgT_f = ar.take<float>(static_cast<size_t>(HW) * F);
    }
    avgT_c = ar.take<float>(static_cast<size_t>(HW) * C);
    avgT_e = ar.take<float>(static_cast<size_t>(HW) * C);
    partial = ar.take<float>(static_cast<size_t>(chunks_f > chunks_c ? chunks_f : chunks_c) * (F > C ? F : C));
    stage = ar.take<float>(32 * static_cast<size_t>(F > C ? F : C));
    tokpart = ar.take<float>(tok_blocks);
  };
  Arena dry;
  dry.dry = true;
  carve(dry);
  SVB_TRY(ensure_arena(h, dry.off));
  h->arena.off = 0; h->arena.dry = false; h->gradbuf = nullptr;
  carve(h->arena);
  const bf16* Xp = zx ? static_cast<const bf16*>(x->x) : X;
  const bf16* Gp = zg ? static_cast<const bf16*>(grad) : G;
  if (!zx) SVB_TRY(pack_acts(st, x, X));
  if (!zg) SVB_TRY(pack_acts(st, &gx, G));
  (prep_encoder_kernel<<<cdiv(F, 8), 256, 0, st>>>(p->w_enc, p->b_enc, p->b_dec, Web, fold, nullptr, F, C), svb::count_launch());
  (convert_kernel<float, bf16><<<grid_for(FC), 256, 0, st>>>(p->w_dec, Wdb, FC), svb::count_launch());
  if (!fused) (transpose_f32_kernel<<<dim3(cdiv(HW, 32), cdiv(F, 32)), dim3(32, 8), 0, st>>>(enc_avg, avgT_f, F, HW), svb::count_launch());
  (transpose_f32_kernel<<<dim3(cdiv(HW, 32), cdiv(C, 32)), dim3(32, 8), 0, st>>>(x_avg, avgT_c, C, HW), svb::count_launch());
  (transpose_f32_kernel<<<dim3(cdiv(HW, 32), cdiv(C, 32)), dim3(32, 8), 0, st>>>(err_avg, avgT_e, C, HW), svb::count_launch());
  SVB_LAUNCH_CHECK("node_ie prep");
  if (fused) {
    SVB_GEMM(launch_fused_node_ie(st, Xp, Gp, Web, Wdb, fold, enc_avg, Ti, C, F, HW, ie_part, q_part, h->sms), "fused node-IE");
    if (ie_features) SVB_TRY(reduce_rows(st, ie_part, 2 * ie_slots, F, scale, stage, ie_features));
    if (ie_neurons)
      SVB_TRY(launch_ie_channelwise<bf16>(st, h->sms, Xp, Gp, avgT_c, T, HW, C, scale, partial, chunks_c, stage, ie_neurons));
    if (ie_error) {
      (ie_error_tokens_fused_kernel<<<tok_blocks, 256, 0, st>>>(Xp, Gp, avgT_e, p->b_dec, q_part, q_rows, T, C, HW, tokpart), svb::count_launch());
      (reduce_flat_kernel<<<1, 1024, 0, st>>>(tokpart, static_cast<size_t>(tok_blocks), scale, ie_error), svb::count_launch());
      SVB_LAUNCH_CHECK("ie_error (fused)");
    }
    return 0;
  }
  // a = SAE_enc(x)
  EpiEnc::Params e1{};
  e1.bias = fold; e1.e_bf16 = E; e1.words = (F + 31) / 32;
  if (make_store_tmap_bf16_chunk(&e1.tm_e, E, Ti, F, F)) return fail(SVB_ERR_TMAP, "tensor map for E");
  SVB_GEMM((launch_gemm_s<false, false, EpiEnc>(st, Xp, C, Web, C, Ti, F, C, 1, e1)), "enc");
  // DIFF = dec - x = -(sae error)
  EpiDec::Params e2{};
  e2.bias = p->b_dec; e2.x = Xp; e2.diff_bf16 = DIFF;
  if (make_store_tmap_bf16(&e2.tm_diff, DIFF, Ti, C, C)) return fail(SVB_ERR_TMAP, "tensor map for DIFF");
  SVB_GEMM((launch_gemm_s<false, false, EpiDec>(st, E, F, Wdb, F, Ti, C, F, 1, e2)), "dec");
  // enc.grad = g W_dec   (nnsight_intervention_check.py:194-195)
  EpiStore::Params e3;
  make_store_params(&e3, GE, F, nullptr, 1.f, 0, 1, Ti, F);
  SVB_GEMM((launch_gemm_s<false, true, EpiStore>(st, Gp, C, Wdb, F, Ti, F, C, 1, e3)), "g W_dec");
  if (ie_features)
    SVB_TRY(launch_ie_channelwise<bf16>(st, h->sms, E, GE, avgT_f, T, HW, F, scale, partial, chunks_f, stage, ie_features));
  if (ie_neurons)
    SVB_TRY(launch_ie_channelwise<bf16>(st, h->sms, Xp, Gp, avgT_c, T, HW, C, scale, partial, chunks_c, stage, ie_neurons));
  if (ie_error) {
    (ie_allchannels_tokens_kernel<<<tok_blocks, 256, 0, st>>>(DIFF, Gp, avgT_e, T, C, HW, -1.f, tokpart), svb::count_launch());
    (reduce_flat_kernel<<<1, 1024, 0, st>>>(tokpart, static_cast<size_t>(tok_blocks), scale, ie_error), svb::count_launch());
    SVB_LAUNCH_CHECK("ie_error");
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------- generic GEMM
namespace {
template <bool AMN, bool BMN>
int gemm_dispatch(svb_handle* h, cudaStream_t st, const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N,
                  int K, void* out, int out_dtype, int64_t ldo, float alpha, const float* bias, int relu) {
  const bool can_split = out_dtype == SVB_F32 && !bias && !relu && ldo == N;
  const int splits = can_split ? planned_splits<256>(M, N, K, 0) : 1;
  if (splits <= 1) {
    EpiStore::Params ep;
    make_store_params(&ep, out, ldo, bias, alpha, relu, out_dtype == SVB_BF16 ? 1 : 0, M, N);
    SVB_GEMM((launch_gemm<256, AMN, BMN, EpiStore>(st, A, lda, B, ldb, M, N, K, 1, ep)), "svb_gemm_bf16");
    return 0;
  }
  const size_t MN = static_cast<size_t>(M) * N;
  Arena dry; dry.dry = true;
  dry.take<float>(static_cast<size_t>(splits) * MN);
  SVB_TRY(ensure_arena(h, dry.off));
  h->arena.off = 0; h->arena.dry = false; h->gradbuf = nullptr;
  float* part = h->arena.take<float>(static_cast<size_t>(splits) * MN);
  EpiPartial::Params ep{part, N, static_cast<long long>(MN)};
  int used = 0;
  SVB_GEMM((launch_gemm<256, AMN, BMN, EpiPartial>(st, A, lda, B, ldb, M, N, K, splits, ep, &used)), "svb_gemm_bf16");
  (sum_splits_kernel<<<grid_for(MN), 256, 0, st>>>(part, used, MN, alpha, static_cast<float*>(out)), svb::count_launch());
  SVB_LAUNCH_CHECK("sum_splits");
  return 0;
}
}  // namespace

extern "C" int svb_gemm_bf16(svb_handle* h, void* stream, const void* A, int32_t a_mn, int64_t lda, const void* B,
                             int32_t b_mn, int64_t ldb, int32_t M, int32_t N, int32_t K, void* out, int32_t out_dtype,
                             int64_t ldo, float alpha, const float* bias, int32_t relu) {
  if (!h || !A || !B || !out) return fail(SVB_ERR_BAD_ARG, "null argument");
  SVB_ON_DEVICE(h);
  if (out_dtype != SVB_F32 && out_dtype != SVB_BF16) return fail(SVB_ERR_BAD_ARG, "bad out dtype");
  if (N % 8 || ldo % 8) return fail(SVB_ERR_UNSUPPORTED, "N and ldo must be multiples of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!a_mn && !b_mn) return gemm_dispatch<false, false>(h, st, A, lda, B, ldb, M, N, K, out, out_dtype, ldo, alpha, bias, relu);
  if (!a_mn && b_mn) return gemm_dispatch<false, true>(h, st, A, lda, B, ldb, M, N, K, out, out_dtype, ldo, alpha, bias, relu);
  if (a_mn && !b_mn) return gemm_dispatch<true, false>(h, st, A, lda, B, ldb, M, N, K, out, out_dtype, ldo, alpha, bias, relu);
  return gemm_dispatch<true, true>(h, st, A, lda, B, ldb, M, N, K, out, out_dtype, ldo, alpha, bias, relu);
}
