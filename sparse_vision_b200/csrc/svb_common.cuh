// Handle, error plumbing and the bump-allocated workspace shared by the svb_* entry points.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "../../include/svb.h"
#include "gemm_host.cuh"
#include "epilogues.cuh"
#include "kernels_misc.cuh"
#include "kernels_ie.cuh"

namespace svb {

inline char* err_buf() {
  static thread_local char buf[512] = "";
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define SVB_CUDA(expr)                                                                                    \
  do {                                                                                                    \
    cudaError_t e_ = (expr);                                                                              \
    if (e_ != cudaSuccess) return svb::fail(SVB_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)
#define SVB_LAUNCH_CHECK(what)                                                                              \
  do {                                                                                                      \
    cudaError_t e_ = cudaGetLastError();                                                                    \
    if (e_ != cudaSuccess) return svb::fail(SVB_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e_)); \
  } while (0)
#define SVB_TRY(expr)                \
  do {                               \
    int rc_ = (expr);                \
    if (rc_ != 0) return rc_;        \
  } while (0)
#define SVB_GEMM(expr, what)                                                             \
  do {                                                                                   \
    int rc_ = (expr);                                                                    \
    if (rc_ != 0) return svb::fail(rc_, "GEMM %s could not be launched (rc=%d)", what, rc_); \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int cdiv(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }
inline int grid_for(size_t n, int threads = 256, int cap = 148 * 16) {
  size_t g = (n + threads - 1) / threads;
  if (g > static_cast<size_t>(cap)) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// Bump allocator over one cudaMalloc'd arena.  A call first measures (dry = true), grows the arena if needed, then
// carves for real; two calls with the same shapes carve identical addresses (step_grads / step_apply rely on it).
struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0, off = 0;
  bool dry = false;
  template <typename T>
  T* take(size_t n) {
    const size_t bytes = align_up(n * sizeof(T), 256);
    T* p = dry ? nullptr : reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
};

}  // namespace svb

namespace svb {
// Optional per-phase timing of the training step with CUDA events on the caller's stream (bench.py's roofline).
constexpr int kProfMaxSteps = 128, kProfMaxMarks = 12;
struct Profiler {
  bool on = false;
  int step = -1;                       // ring slot of the step being recorded
  long long steps_recorded = 0;
  cudaEvent_t ev[kProfMaxSteps][kProfMaxMarks];
  int marks[kProfMaxSteps];
  bool created = false;
};
}  // namespace svb

struct svb_comm;  // svb_comm.cu

struct svb_handle {
  int device = 0;
  int sms = 0;
  svb::Arena arena;
  svb::Profiler prof;
  // description of the flat reduction buffer of the last *_step_grads call
  float* gradbuf = nullptr;
  int64_t sum_elems = 0, max_elems = 0;
  int32_t step_flags = 0;           // SVB_STEP_* bits of the last *_step_grads call (svb_last_step_flags)
  int64_t early_elems = 0;          // leading elements that are final when `comm` is released (0: no early bucket)
  // data-parallel overlap: an optional caller stream that is made to wait for the early bucket (svb_set_comm_stream)
  cudaStream_t comm = nullptr;
  cudaEvent_t ev_early = nullptr;
  svb_comm* comm_ctx = nullptr;     // peer-memory exchange buffer (svb_comm_alloc)
  // internal side stream: small kernels that only feed the end of the step run beside the GEMMs (fork / join below)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  float* coef_dev = nullptr;        // {lr / bc1, 1 / sqrt(bc2)} of the running step when the step count lives on the device
};

namespace svb {
float* comm_flat_or(svb_handle* h, float* arena_flat, size_t need_floats);

// Every entry point runs on the device its handle was created on, whatever the caller's current device is (tensors on
// cuda:1 while the default device is cuda:0): switch for the duration of the call, restore on the way out.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; prev = -1; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    if (prev == dev) prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define SVB_ON_DEVICE(h)                           \
  svb::DeviceGuard device_guard_((h)->device);     \
  if (!device_guard_.ok) return svb::fail(SVB_ERR_CUDA, "could not switch to device %d of the handle", (h)->device)
}

namespace svb {

inline void prof_begin_step(svb_handle* h) {
  Profiler& p = h->prof;
  if (!p.on) return;
  p.step = static_cast<int>(p.steps_recorded % kProfMaxSteps);
  p.marks[p.step] = 0;
  ++p.steps_recorded;
}
inline void prof_mark(svb_handle* h, cudaStream_t st, int idx) {
  Profiler& p = h->prof;
  if (!p.on || p.step < 0 || idx >= kProfMaxMarks) return;
  cudaEventRecord(p.ev[p.step][idx], st);
  if (idx + 1 > p.marks[p.step]) p.marks[p.step] = idx + 1;
}

inline int ensure_arena(svb_handle* h, size_t need) {
  if (need <= h->arena.cap) return 0;
  if (h->arena.base) {
    cudaError_t e = cudaFree(h->arena.base);  // synchronises with any work still using the old arena
    h->arena.base = nullptr;
    h->arena.cap = 0;
    if (e != cudaSuccess) return fail(SVB_ERR_CUDA, "cudaFree failed: %s", cudaGetErrorString(e));
  }
  const size_t want = align_up(need + need / 8, 1 << 20);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(SVB_ERR_NOMEM, "workspace of %zu bytes could not be allocated: %s", want, cudaGetErrorString(e));
  }
  h->arena.base = static_cast<uint8_t*>(p);
  h->arena.cap = want;
  return 0;
}

// Fork / join of the handle's side stream around the caller's stream `st`.  The GEMMs leave 4 of the 148 SMs idle
// (144-CTA grids) and are bandwidth / tensor bound, so the latency-bound helper kernels (statistics folds, column-sum
// reduction, gradient assembly, weight prologue) cost nothing when they run beside them.  Everything forked is joined
// again before the call returns, so the caller still sees plain stream semantics (and may capture the call in a graph).
inline int side_fork(svb_handle* h, cudaStream_t st) {
  if (!h->side) {
    SVB_CUDA(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    SVB_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    SVB_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  }
  SVB_CUDA(cudaEventRecord(h->ev_fork, st));
  SVB_CUDA(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
  return 0;
}
inline int side_join(svb_handle* h, cudaStream_t st) {
  SVB_CUDA(cudaEventRecord(h->ev_join, h->side));
  SVB_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
  return 0;
}

// Releases the caller's communication stream once everything enqueued on `st` so far has finished.
inline int release_comm_stream(svb_handle* h, cudaStream_t st) {
  if (!h->comm) return 0;
  if (!h->ev_early && cudaEventCreateWithFlags(&h->ev_early, cudaEventDisableTiming) != cudaSuccess)
    return fail(SVB_ERR_CUDA, "cudaEventCreate failed");
  SVB_CUDA(cudaEventRecord(h->ev_early, st));
  SVB_CUDA(cudaStreamWaitEvent(h->comm, h->ev_early, 0));
  return 0;
}

inline int check_acts(const svb_acts* x) {
  if (!x || !x->x) return fail(SVB_ERR_BAD_ARG, "activations pointer is null");
  if (x->dtype != SVB_F32 && x->dtype != SVB_BF16) return fail(SVB_ERR_BAD_ARG, "bad activation dtype %d", x->dtype);
  if (x->layout != SVB_TOKENS && x->layout != SVB_NCHW) return fail(SVB_ERR_BAD_ARG, "bad layout %d", x->layout);
  if (x->n_images <= 0 || x->hw <= 0 || x->C <= 0) return fail(SVB_ERR_BAD_ARG, "empty activation batch");
  if (x->C % 8) return fail(SVB_ERR_UNSUPPORTED, "act_size C=%d must be a multiple of 8", x->C);
  if (x->n_images * static_cast<long long>(x->hw) >= (1LL << 31)) return fail(SVB_ERR_UNSUPPORTED, "too many tokens");
  return 0;
}

// Token-major bf16 view of the SAE input: zero-copy when the caller already provides it, else packed into `buf`.
inline bool acts_are_bf16_tokens(const svb_acts* x) {
  return x->dtype == SVB_BF16 && (x->layout == SVB_TOKENS || x->hw == 1) &&
         (reinterpret_cast<uintptr_t>(x->x) & 15) == 0;
}
// Widest per-row access (in elements) that an NCHW tensor with HW positions per row and base pointer p allows.
inline int nchw_vec(const void* p, int hw, int elem_bytes) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  if (hw % 8 == 0 && (a & 15) == 0) return 8;
  if (hw % 4 == 0 && (a % (4 * elem_bytes)) == 0) return 4;
  return 1;
}
// slab_major: write [ceil(C/64)][T][64] instead of [T, C] (NCHW inputs only; see gemm_host.cuh)
inline int pack_acts(cudaStream_t st, const svb_acts* x, bf16* buf, bool slab_major = false, float* xpart = nullptr) {
  const long long T = x->n_images * static_cast<long long>(x->hw);
  if (slab_major && (x->layout != SVB_NCHW || x->hw == 1)) return fail(SVB_ERR_BAD_ARG, "slab-major pack needs NCHW input");
  const long long slab_rows = slab_major ? T : 0;
  if (x->layout == SVB_TOKENS || x->hw == 1) {
    const size_t n = static_cast<size_t>(T) * x->C;
    if (x->dtype == SVB_F32)
      (convert_kernel<float, bf16><<<grid_for(n), 256, 0, st>>>(static_cast<const float*>(x->x), buf, n), svb::count_launch());
    else
      (convert_kernel<bf16, bf16><<<grid_for(n), 256, 0, st>>>(static_cast<const bf16*>(x->x), buf, n), svb::count_launch());
  } else {
    if (x->n_images > 65535) return fail(SVB_ERR_UNSUPPORTED, "more than 65535 images per call");
    const dim3 grid(cdiv(x->hw, 64), cdiv(x->C, 64), static_cast<unsigned>(x->n_images));
    const int vec = nchw_vec(x->x, x->hw, x->dtype == SVB_F32 ? 4 : 2);
#define SVB_PACK(T, V) (pack_nchw_tile_kernel<T, V><<<grid, 256, 0, st>>>(static_cast<const T*>(x->x), buf, x->C, x->hw, slab_rows, xpart), svb::count_launch())
    if (x->dtype == SVB_F32) { if (vec >= 4) SVB_PACK(float, 4); else SVB_PACK(float, 1); }
    else { if (vec == 8) SVB_PACK(bf16, 8); else if (vec == 4) SVB_PACK(bf16, 4); else SVB_PACK(bf16, 1); }
#undef SVB_PACK
  }
  SVB_LAUNCH_CHECK("pack_acts");
  return 0;
}
inline int unpack_to(cudaStream_t st, const bf16* tok, long long n_images, int hw, int C, void* out, int out_dtype,
                     int out_layout) {
  const size_t n = static_cast<size_t>(n_images) * hw * C;
  if (out_layout == SVB_TOKENS || hw == 1) {
    if (out_dtype == SVB_F32) (convert_kernel<bf16, float><<<grid_for(n), 256, 0, st>>>(tok, static_cast<float*>(out), n), svb::count_launch());
    else (convert_kernel<bf16, bf16><<<grid_for(n), 256, 0, st>>>(tok, static_cast<bf16*>(out), n), svb::count_launch());
  } else {
    if (n_images > 65535) return fail(SVB_ERR_UNSUPPORTED, "more than 65535 images per call");
    dim3 grid(cdiv(hw, 32), cdiv(C, 32), static_cast<unsigned>(n_images));
    dim3 block(32, 8);
    if (out_dtype == SVB_F32)
      (unpack_tokens_to_nchw_kernel<bf16, float><<<grid, block, 0, st>>>(tok, static_cast<float*>(out), C, hw), svb::count_launch());
    else if (hw % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
      (unpack_tokens_bf16_fast_kernel<<<dim3(cdiv(hw, 64), cdiv(C, 64), static_cast<unsigned>(n_images)), 256, 0, st>>>(tok, static_cast<bf16*>(out), C, hw), svb::count_launch());
    else
      (unpack_tokens_to_nchw_kernel<bf16, bf16><<<grid, block, 0, st>>>(tok, static_cast<bf16*>(out), C, hw), svb::count_launch());
  }
  SVB_LAUNCH_CHECK("unpack");
  return 0;
}

// EpiStore parameters; bf16 outputs with a dense pitch get a TMA store map (rows x cols), others store directly.
inline int make_store_params(EpiStore::Params* ep, void* out, long long ld, const float* bias, float alpha, int relu,
                             int out_bf16, long long rows, long long cols) {
  memset(ep, 0, sizeof(*ep));
  ep->out = out; ep->ld = ld; ep->bias = bias; ep->alpha = alpha; ep->relu = relu;
  ep->out_bf16 = out_bf16;
  if (out_bf16 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (ld % 8) == 0) {
    if (make_store_tmap_bf16(&ep->tm, out, rows, cols, ld) == 0) ep->tm_valid = 1;
  }
  return 0;
}

// out[j] = scale * sum_i in[i, j] over R rows with a fixed order; `stage` holds up to 32*N floats.
inline int reduce_rows(cudaStream_t st, const float* in, int R, int N, float scale, float* stage, float* out) {
  int chunks = R >= 256 ? 32 : 1;
  if (chunks > 1) {
    (reduce_rows_kernel<<<dim3(cdiv(N, 32), chunks), 256, 0, st>>>(in, stage, R, N, static_cast<size_t>(N), 1.f), svb::count_launch());
    (reduce_rows_kernel<<<dim3(cdiv(N, 32), 1), 256, 0, st>>>(stage, out, chunks, N, static_cast<size_t>(N), scale), svb::count_launch());
  } else {
    (reduce_rows_kernel<<<dim3(cdiv(N, 32), 1), 256, 0, st>>>(in, out, R, N, static_cast<size_t>(N), scale), svb::count_launch());
  }
  SVB_LAUNCH_CHECK("reduce_rows");
  return 0;
}

// Channel statistics of one batch (see channel_stats_kernel).  st holds stats_elems(...) floats.
constexpr int kStatRows = 392;
inline size_t stats_elems(long long n_img, int hw, long long T, int C) {
  const long long imgs = hw > 1 ? n_img : 1, rows = hw > 1 ? hw : T;
  const int chunks = cdiv(rows, kStatRows);
  const int chunks64 = hw > 1 ? cdiv(hw, 64) : 1;  // post_dec_nchw_kernel may split down to one 64-position tile per chunk
  return static_cast<size_t>(imgs) * (chunks > chunks64 ? chunks : chunks64) * 8 * C;
}
inline int run_channel_stats(cudaStream_t st, const bf16* X, const bf16* D, long long n_img, int hw, long long T,
                             int C, float* stbuf, float* chan, float* var_part, float* rowvar) {
  const long long imgs = hw > 1 ? n_img : 1, rows = hw > 1 ? hw : T;
  const int R = cdiv(rows, kStatRows);
  if (imgs > 2147483647LL || R > 65535) return fail(SVB_ERR_UNSUPPORTED, "batch too large for the stats kernel");
  (channel_stats_kernel<<<dim3(static_cast<unsigned>(imgs), R, cdiv(C, 256)), 256, 0, st>>>(X, D, stbuf, C, static_cast<int>(rows), kStatRows), svb::count_launch());
  (channel_stats_finalize_kernel<<<cdiv(C, 8), 256, 0, st>>>(stbuf, chan, var_part, static_cast<int>(imgs), R, C, static_cast<int>(rows)), svb::count_launch());
  if (hw == 1) (row_variance_kernel<<<cdiv(T, 8), 256, 0, st>>>(X, D, rowvar, static_cast<int>(T), C), svb::count_launch());
  SVB_LAUNCH_CHECK("channel_stats");
  return 0;
}

// Fused statistics + NCHW write-back after the decoder GEMM (post_dec_nchw_kernel) when the SAE input is NCHW;
// otherwise the token-major statistics kernel followed by a layout/dtype conversion of d.
//   x: the caller's activations; X / D: token-major bf16 views; dec_out may be null.
// The fused kernel is used when post_dec_fusable() holds; only then may D be slab-major.
inline bool post_dec_fusable(const svb_acts* x, const void* dec_out, int dec_layout) {
  return x->layout == SVB_NCHW && x->hw > 1 && (!dec_out || dec_layout == SVB_NCHW) && x->n_images <= 65535;
}
inline int run_post_dec(cudaStream_t st, const svb_acts* x, const bf16* X, const bf16* D, long long T, void* dec_out,
                        int dec_dtype, int dec_layout, float* stbuf, float* chan, float* var_part, float* rowvar,
                        bool d_slab = false) {
  const int C = x->C, hw = x->hw;
  const long long slab_rows = d_slab ? T : 0;
  if (post_dec_fusable(x, dec_out, dec_layout)) {
    int vec = nchw_vec(x->x, hw, x->dtype == SVB_F32 ? 4 : 2);
    if (dec_out) {
      const int vo = nchw_vec(dec_out, hw, dec_dtype == SVB_F32 ? 4 : 2);
      if (vo < vec) vec = vo;
    }
    // enough blocks for ~4 per SM: split the HW range when images x channel groups alone are too few
    const int tiles = cdiv(hw, 64);
    int R = 1;
    while (R < tiles && static_cast<long long>(cdiv(C, 64)) * x->n_images * R < 4LL * device_sm_count()) ++R;
    const int tpc = cdiv(tiles, R);
    R = cdiv(tiles, tpc);
    if (stats_elems(x->n_images, hw, T, C) < static_cast<size_t>(x->n_images) * R * 8 * C) { R = 1; }
    const int tiles_per_chunk = R == 1 ? tiles : tpc;
    const dim3 grid(cdiv(C, 64), static_cast<unsigned>(x->n_images), R);
#define SVB_POST(TI, TO, V)                                                                                          \
  (post_dec_nchw_kernel<TI, TO, V><<<grid, 256, 0, st>>>(D, static_cast<const TI*>(x->x), static_cast<TO*>(dec_out), \
                                                          stbuf, C, hw, tiles_per_chunk, slab_rows), svb::count_launch())
#define SVB_POST_V(TI, TO)                                     \
  do {                                                         \
    if (vec == 8) SVB_POST(TI, TO, 8);                         \
    else if (vec == 4) SVB_POST(TI, TO, 4);                    \
    else SVB_POST(TI, TO, 1);                                  \
  } while (0)
    const bool out_f32 = dec_out && dec_dtype == SVB_F32;
    if (x->dtype == SVB_F32) { if (out_f32) SVB_POST_V(float, float); else SVB_POST_V(float, bf16); }
    else { if (out_f32) SVB_POST_V(bf16, float); else SVB_POST_V(bf16, bf16); }
#undef SVB_POST_V
#undef SVB_POST
    (channel_stats_finalize_kernel<<<cdiv(C, 8), 256, 0, st>>>(stbuf, chan, var_part, static_cast<int>(x->n_images), R, C, hw), svb::count_launch());
    SVB_LAUNCH_CHECK("post_dec");
    return 0;
  }
  if (d_slab) return fail(SVB_ERR_BAD_ARG, "slab-major decoder output needs the fused post-decoder pass");
  SVB_TRY(run_channel_stats(st, X, D, x->n_images, hw, T, C, stbuf, chan, var_part, rowvar));
  if (dec_out) SVB_TRY(unpack_to(st, D, x->n_images, hw, C, dec_out, dec_dtype, dec_layout));
  return 0;
}

inline int run_prep_step(cudaStream_t st, PrepArgs a) {
  a.nb_enc = cdiv(a.F, 8);
  a.nb_dec = grid_for(static_cast<size_t>(a.F) * a.C / 4, 256, 512);
  a.nb_zero = a.n_zero ? grid_for(a.n_zero, 256, 64) : 0;
  a.nb_exp = a.exp_r ? cdiv(a.F, 256) : 0;
  (prep_step_kernel<<<a.nb_enc + a.nb_dec + a.nb_zero + a.nb_exp, 256, 0, st>>>(a), svb::count_launch());
  SVB_LAUNCH_CHECK("prep_step");
  return 0;
}
// part: 1 = everything that depends on the encoder-side GEMM only (g_wenc, g_benc, the vecmat partials for g_bdec) plus
// the activity counts, 2 = the decoder weight gradient, 3 = both.  The split lets the encoder-side gradients leave
// for the data-parallel all-reduce while the decoder weight-gradient GEMM is still running.
inline int run_assemble(cudaStream_t st, AssembleArgs a, int part = 3) {
  const int nbw = grid_for(static_cast<size_t>(a.F) * a.C / 4, 256, 1024);
  a.nb_wd = (part & 2) ? nbw : 0;
  a.nb_w = (part & 1) ? nbw : 0;
  a.nb_b = (part & 1) ? cdiv(a.F, 256) : 0;
  a.nb_vm = (part & 1) ? a.vm_chunks * cdiv(a.C, 256) : 0;
  a.nb_cnt = (part & 1) ? a.words : 0;
  a.nb_img = (part & 1) ? cdiv(a.n_img, 8) : 0;
  const int grid = a.nb_wd + a.nb_w + a.nb_b + a.nb_vm + a.nb_cnt + a.nb_img;
  if (grid == 0) return 0;
  (assemble_grads_kernel<<<grid, 256, 0, st>>>(a), svb::count_launch());
  SVB_LAUNCH_CHECK("assemble_grads");
  return 0;
}
inline int run_adam_multi(cudaStream_t st, const AdamSeg* segs, int nseg, const AdamCoef& k) {
  AdamMultiArgs a;
  memset(&a, 0, sizeof(a));
  a.nseg = nseg;
  int blk = 0;
  for (int i = 0; i < nseg; ++i) {
    a.seg[i] = segs[i];
    a.blk0[i] = blk;
    blk += grid_for(static_cast<size_t>(segs[i].n), 256, 1024);
  }
  for (int i = nseg; i < 7; ++i) a.blk0[i] = blk;
  (adam_multi_kernel<<<blk, 256, 0, st>>>(a, k), svb::count_launch());
  SVB_LAUNCH_CHECK("adam_multi");
  return 0;
}
inline void launch_cadam(cudaStream_t st, float* w, float* g, float* m, float* v, int C, int F, const AdamCoef& k) {
  (constrained_adam_decoder_kernel<<<cdiv(F, kCadamCols), kCadamCols * kCadamRows, 0, st>>>(w, g, m, v, C, F, k), svb::count_launch());
}
// Scatter of the channel-major copy of d (EpiDecNchw out_kind 4) into the caller's NCHW tensor
inline int run_cmajor_to_nchw(cudaStream_t st, const bf16* dt, void* out, int out_dtype, int C, int hw, long long T,
                              long long ld) {
  if (out_dtype == SVB_F32 && hw % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    (cmajor_to_nchw_kernel<float, 4><<<dim3(static_cast<unsigned>(cdiv(T, 4096)), C), 256, 0, st>>>(dt, static_cast<float*>(out), C, hw, T, ld), svb::count_launch());
  } else if (out_dtype == SVB_F32) {
    (cmajor_to_nchw_kernel<float, 1><<<dim3(static_cast<unsigned>(cdiv(T, 1024)), C), 256, 0, st>>>(dt, static_cast<float*>(out), C, hw, T, ld), svb::count_launch());
  } else if (hw % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0) {
    (cmajor_to_nchw_kernel<bf16, 4><<<dim3(static_cast<unsigned>(cdiv(T, 4096)), C), 256, 0, st>>>(dt, static_cast<bf16*>(out), C, hw, T, ld), svb::count_launch());
  } else {
    (cmajor_to_nchw_kernel<bf16, 1><<<dim3(static_cast<unsigned>(cdiv(T, 1024)), C), 256, 0, st>>>(dt, static_cast<bf16*>(out), C, hw, T, ld), svb::count_launch());
  }
  SVB_LAUNCH_CHECK("cmajor_to_nchw");
  return 0;
}
// Rows of EpiDPre's per-CTA column-sum partials for a B-stationary launch (mirrors launch_gemm's grid choice).
inline int bstat_groups(int sms, int tiles_n, int tiles_m) {
  int groups = sms / tiles_n;
  if (groups > tiles_m) groups = tiles_m;
  return groups;
}

inline AdamCoef adam_coef(const svb_opt_config* o) {
  AdamCoef k;
  k.dev = nullptr;
  const double bc1 = 1.0 - pow(static_cast<double>(o->beta1), o->step);
  const double bc2 = 1.0 - pow(static_cast<double>(o->beta2), o->step);
  k.lr_over_bc1 = static_cast<float>(static_cast<double>(o->lr) / bc1);
  k.inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(bc2));
  k.beta2 = static_cast<float>(o->beta2);
  k.omb1 = static_cast<float>(1.0 - o->beta1);   // rounded once from the double, as torch does with its Python scalars
  k.omb2 = static_cast<float>(1.0 - o->beta2);
  k.eps = static_cast<float>(o->eps);
  return k;
}

}  // namespace svb

namespace svb {
// Adam coefficients of a call: from the host step count, or (svb_opt_config::step_dev) from a device counter that a
// one-thread kernel on `st` increments first -- everything enqueued after it on `st` (and on streams forked from it)
// sees the new values.
inline int adam_coef_for(svb_handle* h, cudaStream_t st, const svb_opt_config* o, AdamCoef* k) {
  if (!o->step_dev) {
    if (o->step < 1) return fail(SVB_ERR_BAD_ARG, "Adam step must be >= 1");
    *k = adam_coef(o);
    return 0;
  }
  svb_opt_config tmp = *o;
  tmp.step = 1;
  *k = adam_coef(&tmp);
  (adam_step_coef_kernel<<<1, 1, 0, st>>>(o->step_dev, o->lr, o->beta1, o->beta2, h->coef_dev), svb::count_launch());
  SVB_LAUNCH_CHECK("adam_step_coef");
  k->dev = h->coef_dev;
  return 0;
}
}  // namespace svb
