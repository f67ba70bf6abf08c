/* svb.h — C ABI of libsvb.so, the B200 (sm_100a) implementation of sparse-vision's SAE training-and-attribution
 * hot path.  The reference (jasper3100/sparse-vision) is pure Python and has no FFI; these entry points sit UNDER
 * its Python module API and each one names the reference code it replaces (paths relative to the reference root).
 * The ctypes binding a maintainer would add is shown in INTEGRATION.md and lives in sparse_vision_b200/_lib.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the library never frees or retains caller
 *     memory beyond the call; tensors are dense row-major;
 *   - `stream` is a cudaStream_t passed as void* (PyTorch: torch.cuda.current_stream().cuda_stream); all work is
 *     enqueued on it and nothing synchronises with the host;
 *   - return value: 0 on success, negative svb_status otherwise; svb_last_error() gives a message (thread-local);
 *   - parameters and Adam state are fp32 and are updated IN PLACE (the reference indexes them afterwards);
 *   - there is no CPU fallback: a missing device or an unsupported shape is an error.
 *   - shape limits: C % 8 == 0 and F % 8 == 0 (TMA pitch), T = n_images * hw < 2^31.
 */
#ifndef SVB_H_
#define SVB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svb_handle svb_handle; /* opaque; owns workspaces for ONE device; not thread-safe */

enum svb_status {
  SVB_OK = 0,
  SVB_ERR_DRIVER = -1,      /* cuTensorMapEncodeTiled entry point unavailable */
  SVB_ERR_BAD_ARG = -2,     /* null pointer, misaligned pointer, bad enum, bad shape */
  SVB_ERR_TMAP = -3,        /* tensor-map encoding failed */
  SVB_ERR_CUDA = -4,        /* a CUDA runtime call failed */
  SVB_ERR_UNSUPPORTED = -5, /* shape outside the limits above */
  SVB_ERR_NOMEM = -6        /* workspace allocation failed */
};

enum svb_dtype { SVB_F32 = 0, SVB_BF16 = 1 };
enum svb_layout {
  SVB_TOKENS = 0, /* [T, C], C contiguous; token order t = (b*H + h)*W + w  (utils.py:2770-2774 reshape_tensor) */
  SVB_NCHW = 1    /* [B, C, H, W] as the base model produces it */
};
enum svb_optimizer {
  SVB_ADAM = 0,            /* utils.py:85-86  torch.optim.Adam, caller passes betas (0.9, 0.9999) */
  SVB_CONSTRAINED_ADAM = 1 /* utils.py:50-81  project decoder grad, Adam (0.9, 0.999), renormalise columns */
};

const char* svb_last_error(void);
int svb_version(void);
int svb_create(svb_handle** out); /* binds to the calling thread's current CUDA device */
int svb_destroy(svb_handle* h);
/* bytes of device workspace currently held by the handle */
int64_t svb_workspace_bytes(const svb_handle* h);

/* Number of kernels the library has launched in this process (bench.py reports the per-step delta). */
int64_t svb_launch_count(void);

/* Which fused kernels the last svb_sae_step_grads call of this handle used (bit flags), so that per-phase timings
 * can be attributed: SVB_STEP_FUSED_BWD = the dE GEMM, the ReLU mask and the dW_enc GEMM ran as ONE kernel (C <= 256,
 * C % 64 == 0; model_pipeline.py:385 backward of models/sae_mlp.py:49-52) -- its time is reported in the "dE_gemm"
 * phase and the "dWenc_gemm" phase is empty. */
#define SVB_STEP_FUSED_BWD 1
int32_t svb_last_step_flags(const svb_handle* h);

/* Process-wide kernel-selection switches, for A/B measurements and for the tests that pin the fused / paired kernels
 * against the plain ones (results agree within the parity tolerances whatever the values).  Defaults may be preset
 * through the environment variable named beside each key.  svb_get_tuning returns -1 for an unknown key. */
enum svb_tuning_key {
  SVB_TUNE_FUSED_BWD = 0,    /* 1: dE GEMM -> ReLU mask -> dW_enc GEMM as one kernel when C <= 256, C % 64 == 0 (SVB_FUSED_BWD) */
  SVB_TUNE_FBW_2CTA = 1,     /* 1: that kernel on SM pairs (cta_group::2) when C % 128 == 0                      (SVB_FBW_2CTA) */
  SVB_TUNE_GEMM_PAIRS = 2,   /* 1: streaming GEMMs with more than one 128-row tile run on SM pairs                 (SVB_GEMM2) */
  SVB_TUNE_ENC_2CTA = 3,     /* 1: B-stationary encoder GEMM on SM pairs (default 0: measured slower)             (SVB_ENC_2CTA) */
  SVB_TUNE_FBW_PREFETCH = 4, /* L2 prefetch distance (token blocks) of the fused backward, default 0              (SVB_FBW_PF) */
  SVB_TUNE_FUSED_IE = 5,     /* 1: svb_node_ie_layer keeps a and G = g W_dec in TMEM (one kernel) when C % 128 == 0, C <= 256 (SVB_FUSED_IE) */
  SVB_TUNE_ENC16 = 6         /* 1: the B-stationary encoder GEMM (C <= 256) runs 16 epilogue warps instead of 8 (default 0: measured 0.221 against 0.212 ms, the staging costs an operand stage) (SVB_ENC16) */
};
int svb_set_tuning(int32_t key, int32_t value);
int32_t svb_get_tuning(int32_t key);

/* Per-phase timing of the SaeMLP training step with CUDA events recorded on the caller's stream between the phases
 * (pack+prep, enc GEMM, dec GEMM, channel stats, dE GEMM, dW_dec GEMM, dW_enc GEMM, gradient assembly, Adam).
 * svb_profile_read synchronises the device and returns the mean milliseconds per phase over the recorded steps
 * (a ring of the last 128).  ms_avg_host is a HOST array. */
int svb_profile_enable(svb_handle* h, int32_t enable);
int svb_profile_read(svb_handle* h, int32_t max_phases, float* ms_avg_host, int32_t* n_phases, int32_t* n_steps);
const char* svb_profile_phase_name(int32_t i);

/* A batch of SAE inputs: the hooked layer's output (model_pipeline.py:368). */
typedef struct svb_acts {
  const void* x;    /* activations */
  int32_t dtype;    /* svb_dtype */
  int32_t layout;   /* svb_layout */
  int64_t n_images; /* B; for 2-D inputs the number of rows */
  int32_t hw;       /* H*W pixels per image; 1 for 2-D inputs */
  int32_t C;        /* act_size */
} svb_acts;

/* SaeMLP parameters, state_dict order (models/sae_mlp.py:26-40). */
typedef struct svb_sae_params {
  float* w_enc; /* encoder.weight [F, C] */
  float* b_enc; /* encoder.bias   [F]    */
  float* w_dec; /* decoder.weight [C, F] */
  float* b_dec; /* decoder.bias   [C]    */
  int32_t F;    /* hidden_size = C * expansion_factor */
} svb_sae_params;

/* GatedSae parameters, state_dict order (models/gated_sae.py:11-26). */
typedef struct svb_gated_params {
  float* w_gate; /* [F, C] */
  float* b_gate; /* [F] */
  float* b_mag;  /* [F] */
  float* r_mag;  /* [F] */
  float* w_dec;  /* decoder.weight [C, F] */
  float* b_dec;  /* decoder.bias   [C] */
  int32_t F;
} svb_gated_params;

/* Adam moments for the same tensors, same order (torch.optim.Adam state 'exp_avg' / 'exp_avg_sq'). */
typedef struct svb_adam_state {
  float* m[6];
  float* v[6];
} svb_adam_state;

typedef struct svb_opt_config {
  int32_t optimizer; /* svb_optimizer */
  int32_t step;      /* Adam step count AFTER this update (1 on the first step) */
  double lr, beta1, beta2, eps; /* doubles, like the Python floats torch.optim.Adam holds: 1 - beta2 = 1e-4 must not
                                 * pick up the 1.7e-4 relative error of a float32 0.9999 */
  int32_t* step_dev; /* NULL, or a DEVICE int32 holding the number of steps taken so far: the call increments it on the
                      * device and derives the bias corrections from it (`step` is then ignored).  This is what lets a
                      * whole training batch be captured in a CUDA graph and replayed: nothing step-dependent is baked
                      * into the launch parameters. */
} svb_opt_config;

/* Scalars of one step, device float[SVB_STATS_LEN] (read them with one D2H copy per logging interval instead of
 * the reference's six .item() syncs, model_pipeline.py:394-399). */
enum svb_stat {
  SVB_STAT_LOSS = 0, /* rec + lambda*l1 (+ aux)           utils.py:2470,2473 */
  SVB_STAT_REC = 1,  /* mean (d-x)^2                      sparse_loss.py:35 */
  SVB_STAT_L1 = 2,   /* mean |enc| (gated: |relu_pi|)     sparse_loss.py:41,71 */
  SVB_STAT_NRMSE = 3,
  SVB_STAT_RMSE = 4, /* sparse_loss.py:4-21 */
  SVB_STAT_AUX = 5,  /* gated aux mse, else 0             sparse_loss.py:72 */
  SVB_STAT_VAR_EXPL = 6, /* utils.py:2012-2030 */
  SVB_STAT_SPARSITY = 7, /* utils.py:2063-2067 */
  SVB_STAT_N_DEAD = 8,   /* number of units with no activity in this batch */
  SVB_STATS_LEN = 16
};

/* Per-step outputs of the activity bookkeeping (utils.py:2032-2069 measure_inactive_units). */
typedef struct svb_activity_out {
  uint8_t* dead;     /* [F] 1 iff the unit was inactive for every image of the batch; may be NULL */
  float* freq;       /* [F] fraction of images in which the unit fired; may be NULL */
  int32_t* n_active; /* [n_images] active units per image; may be NULL */
} svb_activity_out;

/* ---------------------------------------------------------------------------------------------------------------
 * SaeMLP forward — models/sae_mlp.py:42-53 (4-D branch: pixels as tokens).  Outputs are token-major 2-D like the
 * reference's return values; any of them may be NULL.  enc / pre: [T, F]; dec: [T, C].
 */
typedef struct svb_sae_forward_out {
  void* enc;  int32_t enc_dtype; /* svb_dtype */
  float* pre;                    /* fp32 */
  void* dec;  int32_t dec_dtype;
} svb_sae_forward_out;
int svb_sae_forward(svb_handle* h, void* stream, const svb_acts* x, const svb_sae_params* p,
                    const svb_sae_forward_out* out);

/* ---------------------------------------------------------------------------------------------------------------
 * One SaeMLP training step — the train branch of ModelPipeline.hook (model_pipeline.py:363-432):
 * sae_inference_and_loss (utils.py:2448-2482) -> loss.backward() -> optimizer.step() (utils.py:50-97) ->
 * measure_inactive_units (utils.py:2032-2069) -> variance_explained (utils.py:2012-2030).
 *   dec_out: decoder output in the layout/dtype of `dec_layout`/`dec_dtype` (what the hook returns), or NULL.
 *   global_tokens: tokens over ALL data-parallel ranks (loss means are global); 0 means "this batch only".
 * svb_sae_step_grads + svb_sae_step_apply are the two halves used for data parallelism: the caller all-reduces
 * (SUM) the flat buffer from svb_sae_grad_buffer() between them.  svb_sae_train_step runs both.
 */
typedef struct svb_train_out {
  void* dec_out; int32_t dec_dtype; int32_t dec_layout;
  float* stats;              /* device float[SVB_STATS_LEN] or NULL */
  svb_activity_out activity;
} svb_train_out;

int svb_sae_train_step(svb_handle* h, void* stream, const svb_acts* x, const svb_sae_params* p,
                       const svb_adam_state* adam, const svb_opt_config* opt, float lambda_sparse,
                       int32_t expansion_factor, const svb_train_out* out);
int svb_sae_step_grads(svb_handle* h, void* stream, const svb_acts* x, const svb_sae_params* p, float lambda_sparse,
                       int64_t global_tokens, const svb_train_out* out);
int svb_sae_step_apply(svb_handle* h, void* stream, const svb_acts* x, const svb_sae_params* p,
                       const svb_adam_state* adam, const svb_opt_config* opt, float lambda_sparse,
                       int32_t expansion_factor, int64_t global_tokens, int64_t global_images,
                       const svb_train_out* out);
/* Flat fp32 reduction buffer of the last svb_*_step_grads call: gradients in state_dict order followed by the loss
 * partial sums.  *sum_elems (SUM all-reduce) then *max_elems (MAX all-reduce) elements. */
int svb_sae_grad_buffer(svb_handle* h, float** buf, int64_t* sum_elems, int64_t* max_elems);

/* Data-parallel overlap (the reference has no distributed code).  With a communication stream
 * set, svb_*_step_grads makes that stream wait until the LEADING svb_grad_early_elems() elements of the flat buffer
 * (the encoder-side gradients, [gW_enc | gb_enc] resp. [gW_gate | gb_gate | gb_mag]) are final; the caller
 * all-reduces them on it while the decoder weight-gradient GEMM is still running on `stream`, and the rest of the
 * SUM section afterwards.  Pass NULL to switch the early release off (the default). */
int svb_set_comm_stream(svb_handle* h, void* stream);
int svb_grad_early_elems(svb_handle* h, int64_t* elems);

/* Data-parallel exchange in peer memory (one node, <= 8 ranks).  svb_comm_alloc creates the
 * exchange buffer (capacity n_floats >= sum_elems + max_elems of the steps that will use it) and returns its 64-byte
 * CUDA IPC handle; the caller gathers the handles of all ranks (e.g. torch.distributed.all_gather_object) and passes
 * them, in rank order, to svb_comm_connect.  From then on svb_*_step_grads builds its flat buffer there and
 * svb_comm_allreduce reduces it in place on every rank with ONE kernel over NVLink (SUM section and MAX section,
 * fixed rank order => bit-identical results everywhere).  All ranks must call it once per step, in step order. */
int svb_comm_alloc(svb_handle* h, int64_t n_floats, void* ipc_handle_out_host);
int svb_comm_connect(svb_handle* h, int32_t rank, int32_t world, const void* ipc_handles_host);
/* NVLS variant: the caller owns a SYMMETRIC region of svb_comm_region_bytes(n_floats) zeroed bytes on every rank (same
 * size everywhere), mapped into this process for all ranks (region_ptrs_host[r], r = rank: the local pointer), plus --
 * optionally -- a multicast mapping of it (NULL: plain peer loads / stores as with svb_comm_connect).  With a multicast
 * pointer the SUM section is reduced by the NVSwitch (multimem.ld_reduce / multimem.st).  parallel.py obtains such a
 * region from torch.distributed._symmetric_memory.  The region must outlive the handle's use of it. */
int svb_comm_region_bytes(int64_t n_floats, int64_t* bytes, int64_t* capacity_floats);
int svb_comm_attach(svb_handle* h, int32_t rank, int32_t world, int64_t n_floats, const void* const* region_ptrs_host,
                    void* multicast_ptr);
int svb_comm_capacity(svb_handle* h, int64_t* n_floats);
int svb_comm_allreduce(svb_handle* h, void* stream);
int svb_comm_destroy(svb_handle* h);
/* How long a rank waits in the exchange kernel for a peer (default 120 s, or the environment variable
 * SVB_COMM_TIMEOUT_S when the buffer is allocated).  A peer that never arrives does not hang or trap the GPU: the
 * launch gives up and svb_comm_status reports 1 + that rank (sticky; 0 = all exchanges completed; synchronises). */
int svb_comm_set_timeout(svb_handle* h, double seconds);
int svb_comm_status(svb_handle* h, int32_t* status);

/* GatedSae forward / step — models/gated_sae.py:28-56, losses/sparse_loss.py:68-76. */
typedef struct svb_gated_forward_out {
  void* enc;     int32_t enc_dtype;
  void* dec;     int32_t dec_dtype;
  void* relu_pi; int32_t relu_pi_dtype;
  void* via;     int32_t via_dtype;
} svb_gated_forward_out;
int svb_gated_forward(svb_handle* h, void* stream, const svb_acts* x, const svb_gated_params* p,
                      const svb_gated_forward_out* out);
int svb_gated_train_step(svb_handle* h, void* stream, const svb_acts* x, const svb_gated_params* p,
                         const svb_adam_state* adam, const svb_opt_config* opt, float lambda_sparse,
                         int32_t expansion_factor, const svb_train_out* out);
int svb_gated_step_grads(svb_handle* h, void* stream, const svb_acts* x, const svb_gated_params* p,
                         float lambda_sparse, int64_t global_tokens, const svb_train_out* out);
int svb_gated_step_apply(svb_handle* h, void* stream, const svb_acts* x, const svb_gated_params* p,
                         const svb_adam_state* adam, const svb_opt_config* opt, float lambda_sparse,
                         int32_t expansion_factor, int64_t global_tokens, int64_t global_images,
                         const svb_train_out* out);

/* ---------------------------------------------------------------------------------------------------------------
 * Eval-epoch metrics on the device.
 *
 * svb_spatial_mean    — utils.py:1996-2010 average_over_W_H: mean over the hw positions of every (image, unit).
 *                       `t` is [n_images*hw, F] (SVB_TOKENS, the layout the SAE kernels emit) or [n_images, F, hw]
 *                       (SVB_NCHW), f32 or bf16; out is float[n_images, F].
 * svb_topk_columns    — model_pipeline.py:357-360 (torch.topk(use_output, k, dim=0[, largest=False])) and the merge of
 *                       utils.py:1445-1481 in one call: per column, the k largest (or smallest) of the n0 rows of
 *                       source 0 followed by the n1 rows of source 1 (n1 = 0: plain per-batch top-k), sorted, ties by
 *                       lower position, NaN largest.  idx / files are optional int64 [n, F] payloads gathered along
 *                       (utils.py:1476-1477); a NULL idx stands for the row number.  Outputs are [k, F].
 * svb_histogram_update — utils.py:1934-1963: hist[bins, n_units] += torch.histc(vals[:, unit_idx[u]], bins, mins[u],
 *                       maxs[u]) for every listed unit (integer counts, deterministic).
 */
int svb_spatial_mean(svb_handle* h, void* stream, const void* t, int32_t dtype, int32_t layout, int64_t n_images,
                     int32_t hw, int32_t F, float* out);
/* compute_ie.py:146-207: sums over the images of a batch, per position -- out[per_image] = sum_b t[b*per_image + i] for
 * t = [n_images, per_image] (f32 / bf16; e.g. token-major encoder output with per_image = HW*F, or an NCHW layer output
 * with per_image = C*HW).  Images are added in order (deterministic). */
int svb_image_sum(svb_handle* h, void* stream, const void* t, int32_t dtype, int64_t n_images, int64_t per_image,
                  float* out);
int svb_topk_columns(svb_handle* h, void* stream, const float* vals0, const int64_t* idx0, const int64_t* files0,
                     int32_t n0, const float* vals1, const int64_t* idx1, const int64_t* files1, int32_t n1, int32_t F,
                     int32_t k, int32_t largest, float* out_vals, int64_t* out_idx, int64_t* out_files);
int svb_histogram_update(svb_handle* h, void* stream, const float* vals, int64_t rows, int32_t F,
                         const int64_t* unit_idx, int32_t n_units, const float* mins, const float* maxs, int32_t bins,
                         float* hist);

/* ---------------------------------------------------------------------------------------------------------------
 * Memory-bound layers of the frozen activation producer (utils.py:277-281 builds the base model,
 * model_pipeline.py:445-475, 662-708 run it around the hook).  bf16, NHWC ("channels_last": [n, h, w, c] in memory).
 *
 * svb_maxpool_nhwc      — torch.nn.MaxPool2d(kernel, stride, pad, ceil_mode) of torchvision's GoogLeNet (3/2, 3/1 and
 *                         2/2 windows; padding counts as -inf, NaN propagates); OH / OW are torch's output sizes and
 *                         are checked.  Exact.
 * svb_bias_relu_scatter — what follows every convolution of the model (BasicConv2d: conv -> folded BatchNorm bias ->
 *                         relu) and the torch.cat of an inception block in one pass: out = relu(src + bias) for a dense
 *                         [positions, C] convolution output, channel range [c_begin, c_begin + c_count) written to
 *                         dst[position * dst_channels + dst_offset + (c - c_begin)].  The ranges are consecutive, cover
 *                         C, and every count / offset is a multiple of 8.  Bit-identical to add_ + relu_ on bf16 tensors.
 */
#define SVB_MAX_CHAN_SEGMENTS 4
typedef struct svb_chan_segment {
  void* dst;            /* bf16 [positions, dst_channels] */
  int32_t c_begin;      /* first source channel of the range */
  int32_t c_count;
  int32_t dst_channels; /* row length of dst in channels */
  int32_t dst_offset;   /* first channel of the range inside a dst row */
} svb_chan_segment;
int svb_maxpool_nhwc(svb_handle* h, void* stream, const void* in, int64_t n_images, int32_t H, int32_t W, int32_t C,
                     int32_t kernel, int32_t stride, int32_t pad, int32_t ceil_mode, void* out, int32_t OH, int32_t OW);
int svb_bias_relu_scatter(svb_handle* h, void* stream, const void* src, const void* bias, int64_t positions, int32_t C,
                          const svb_chan_segment* seg, int32_t n_seg, int32_t relu);

/* The differentiable pair for the IE passes (compute_ie.py:270-311 back-propagates the loss through the base model's
 * layers behind the first hooked one).  svb_maxpool_nhwc_argmax also writes, per output element, the window offset
 * kh * kernel + kw of its maximum (one byte; ties and NaN resolved like torch.nn.functional.max_pool2d_with_indices:
 * the first maximum in (kh, kw) order); svb_maxpool_nhwc_backward routes grad_out through those indices:
 * grad_in[n, ih, iw, c] = sum of grad_out over the windows whose maximum sits there (a gather: deterministic). */
int svb_maxpool_nhwc_argmax(svb_handle* h, void* stream, const void* in, int64_t n_images, int32_t H, int32_t W,
                            int32_t C, int32_t kernel, int32_t stride, int32_t pad, int32_t ceil_mode, void* out,
                            uint8_t* argmax, int32_t OH, int32_t OW);
int svb_maxpool_nhwc_backward(svb_handle* h, void* stream, const void* grad_out, const uint8_t* argmax,
                              int64_t n_images, int32_t H, int32_t W, int32_t C, int32_t kernel, int32_t stride,
                              int32_t pad, int32_t OH, int32_t OW, void* grad_in);
/* Backward of svb_bias_relu_scatter for the IE passes: the dense gradient of a convolution output whose relu(. + bias)
 * went to channel range(s) of other tensors.  dst[t, c_begin + c] = y[t, y_offset + c] > 0 ? grad[t, grad_offset + c] : 0
 * (torch's threshold_backward: grad * (result > 0)); `y` holds the forward RESULT of the range, `grad` the gradient
 * that arrived for it.  Ranges consecutive from 0, covering C; everything a multiple of 8; bf16. */
typedef struct svb_grad_segment {
  const void* grad;       /* bf16 [positions, grad_channels] */
  const void* y;          /* bf16 [positions, y_channels] */
  int32_t c_begin, c_count;
  int32_t grad_channels, grad_offset;
  int32_t y_channels, y_offset;
} svb_grad_segment;
int svb_relu_grad_gather(svb_handle* h, void* stream, int64_t positions, int32_t C, const svb_grad_segment* seg,
                         int32_t n_seg, void* dst);
/* GoogLeNet's stem convolution (torchvision googlenet.py conv1: 7x7, stride 2, pad 3, 3 -> 64 channels, 224x224 input)
 * with the folded BatchNorm bias and the ReLU in its epilogue: x bf16 NHWC [n, 224, 224, 3] -> out bf16 NHWC
 * [n, 112, 112, 64] = relu(conv(x, w) + bias), fp32 accumulation, ONE rounding to bf16.  The im2col matrix is never
 * built: with 3 channels a kernel row is 21 contiguous values of the input row (csrc/svb_producer.cu).
 * svb_conv1_pack_weights lays the [64, 3, 7, 7] bf16 weights (element strides given) out once as the kernel's
 * B operand: `packed` is SVB_CONV1_PACKED_ELEMS bf16 values. */
#define SVB_CONV1_PACKED_ELEMS (64 * 168)
int svb_conv1_pack_weights(svb_handle* h, void* stream, const void* w, int64_t stride_o, int64_t stride_i,
                           int64_t stride_h, int64_t stride_w, void* packed);
int svb_conv1_7x7s2_nhwc(svb_handle* h, void* stream, const void* x, int64_t n_images, const void* packed_w,
                         const void* bias, int32_t relu, void* out);

/* ---------------------------------------------------------------------------------------------------------------
 * Optimiser step on caller-provided gradients — utils.py:50-97 (ConstrainedAdam.step / torch.optim.Adam).
 * `decoder_index` is the position of decoder.weight in the lists (projected + renormalised when the optimizer is
 * SVB_CONSTRAINED_ADAM; -1 for none); rows/cols give each tensor's 2-D shape (vectors: rows = 1).
 */
int svb_adam_step(svb_handle* h, void* stream, int32_t n_tensors, float* const* params, const float* const* grads,
                  float* const* m, float* const* v, const int64_t* rows, const int64_t* cols, int32_t decoder_index,
                  const svb_opt_config* opt);

/* Dead-unit re-initialisation scatter — models/sae_mlp.py:133-176.  The caller draws the Kaiming matrices with
 * PyTorch (generator parity) and passes them already rescaled (sae_mlp.py:106-130); this call scatters them into
 * the dead rows / columns, renormalises ALL decoder columns (:138) and zeroes the Adam moments of the dead slices. */
int svb_reinit_dead(svb_handle* h, void* stream, const svb_sae_params* p, int32_t C, const svb_adam_state* adam,
                    const uint8_t* dead, const float* new_w_enc, const float* new_w_dec, float new_b_enc);

/* measure_inactive_units (utils.py:2032-2069) on an existing tensor: [B, F, H, W] (layout NCHW) or [N, F]. */
int svb_measure_inactive(svb_handle* h, void* stream, const void* out_tensor, int32_t dtype, int32_t layout,
                         int64_t n_images, int32_t hw, int32_t F, const svb_activity_out* act);

/* ---------------------------------------------------------------------------------------------------------------
 * Indirect-effect reductions — utils.py:2606-2637 compute_ie_channel_wise and utils.py:2574-2602
 * compute_ie_all_channels.   a, g: [T, F] token-major (dtype f32/bf16); avg: fp32 [F, H, W].
 *   out[f] = scale * sum_t | g[t,f] * (avg[f, t mod HW] - a[t,f]) |      scale = 1/T gives the reference's mean;
 * pass scale = 1 to get partial sums for a data-parallel all-reduce.
 */
int svb_ie_channelwise(svb_handle* h, void* stream, const void* a, const void* g, int32_t dtype, const float* avg,
                       int64_t n_images, int32_t hw, int32_t F, float scale, float* out);
/* err, g: [B, C, H, W] (dtype f32/bf16); avg fp32 [C, H, W]; out[0] = scale * sum_t | sum_c g*(avg - err) |. */
int svb_ie_allchannels(svb_handle* h, void* stream, const void* err, const void* g, int32_t dtype, const float* avg,
                       int64_t n_images, int32_t C, int32_t hw, float scale, float* out);

/* Node-IE for one layer of one batch, fused front to back (compute_ie.py:242-267,442-453 with the identity
 * enc.grad == grad_original @ W_dec, supplementary_files_2/nnsight_intervention_check.py:194-213):
 *   a = SAE_enc(x); G = g W_dec; err = x - dec;
 *   ie_features[F], ie_error[1], ie_neurons[C]  (each scaled by `scale`).  x and g share dtype/layout. */
int svb_node_ie_layer(svb_handle* h, void* stream, const svb_acts* x, const void* grad, const svb_sae_params* p,
                      const float* enc_avg, const float* err_avg, const float* x_avg, float scale,
                      float* ie_features, float* ie_error, float* ie_neurons);

/* Generic bf16 GEMM on the same tcgen05 kernel:  D[M,N] = alpha * A[M,K] * B[N,K]^T (+ bias[N]) (optionally ReLU).
 *   a_mn / b_mn: 0 = K contiguous (memory [M,K] / [N,K] with pitch lda / ldb), 1 = M / N contiguous (memory [K,M] /
 *   [K,N]).  out: fp32 or bf16 with pitch ldo.  With few output tiles and fp32 output (no bias/ReLU) the K loop is
 *   split over the SMs and reduced in a fixed order.  Used by the module-level autograd path (backward of
 *   models/sae_mlp.py and models/gated_sae.py, i.e. what loss.backward() at model_pipeline.py:385 runs). */
int svb_gemm_bf16(svb_handle* h, void* stream, const void* A, int32_t a_mn, int64_t lda, const void* B, int32_t b_mn,
                  int64_t ldb, int32_t M, int32_t N, int32_t K, void* out, int32_t out_dtype, int64_t ldo, float alpha,
                  const float* bias, int32_t relu);

/* Layout helpers (einops rearrange at sae_mlp.py:44 / utils.py:2462-2480), exposed for callers and tests. */
int svb_pack_tokens(svb_handle* h, void* stream, const svb_acts* x, void* out_bf16_tokens);
int svb_unpack_tokens(svb_handle* h, void* stream, const void* tokens, int32_t tokens_dtype, int64_t n_images,
                      int32_t hw, int32_t C, void* out_nchw, int32_t out_dtype);

#ifdef __cplusplus
}
#endif
#endif /* SVB_H_ */
