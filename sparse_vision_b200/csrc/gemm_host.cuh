// Host side of the tcgen05 GEMM: TMA tensor-map encoding (driver entry point fetched at run time, no -lcuda) and
// the launcher that sizes the persistent grid and the split-K slices.
#pragma once
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "gemm_sm100.cuh"

namespace svb {

// Every kernel launch of the library bumps this counter (bench.py reports it as gpu_launches).
inline unsigned long long& launch_counter() {
  static unsigned long long n = 0;
  return n;
}
inline void count_launch() { ++launch_counter(); }

// Process-wide kernel-selection switches (svb_set_tuning / svb_get_tuning in svb.h; keys = SVB_TUNE_* there).  Defaults
// come from the environment once: SVB_FUSED_BWD, SVB_FBW_2CTA, SVB_GEMM2, SVB_ENC_2CTA, SVB_FBW_PF, SVB_FUSED_IE, SVB_ENC16.  They exist for A/B
// measurements and for the tests that pin the fused kernels against the un-fused path; results agree within the
// parity tolerances whatever their values.
enum { kTuneFusedBwd = 0, kTuneFbwTwoCta = 1, kTuneGemmPairs = 2, kTuneEncTwoCta = 3, kTuneFbwPrefetch = 4, kTuneFusedIe = 5, kTuneEnc16 = 6, kTuneCount = 7 };
inline int& tuning(int key) {
  static int v[kTuneCount];
  static const bool init = [] {
    const char* names[kTuneCount] = {"SVB_FUSED_BWD", "SVB_FBW_2CTA", "SVB_GEMM2", "SVB_ENC_2CTA", "SVB_FBW_PF", "SVB_FUSED_IE", "SVB_ENC16"};
    const int defaults[kTuneCount] = {1, 1, 1, 0, 0, 1, 0};
    for (int i = 0; i < kTuneCount; ++i) {
      const char* e = getenv(names[i]);
      v[i] = e ? atoi(e) : defaults[i];
    }
    return true;
  }();
  (void)init;
  return v[key];
}

#ifdef SVB_GEMM_TRACE
inline long long*& gemm_trace_ptr() {  // bring-up only (tools/gemm_selftest.cu): device buffer [grid][4] of wait cycles
  static long long* p = nullptr;
  return p;
}
inline int& gemm_a_skip() {
  static int v = 0;
  return v;
}
#endif

inline PFN_cuTensorMapEncodeTiled_v12000 tmap_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// Row-major bf16 matrix [rows, cols] with row pitch ld (elements); box = box_rows x 64 columns, 128B swizzle.
// Out-of-bounds elements are zero-filled, which is what makes M/N/K tails free.
inline int make_tmap_bf16_2d(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_rows) {
  auto fn = tmap_encode_fn();
  if (!fn) return -1;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15)) return -2;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -3;
}

// Slab-major bf16 matrix: logical [rows, cols] stored as [cols/64][rows][64], i.e. every 64-column slab keeps all its
// rows contiguous (128 bytes per row).  A 128 x 64 operand tile or a 32 x 64 epilogue slab is then ONE contiguous
// block of memory instead of 128-byte pieces at a row pitch of cols*2 bytes, which is what DRAM pages like; rows
// beyond `rows` are clipped / zero-filled exactly as in the row-major case.  cols % 64 != 0: the last slab is stored
// whole and its padding columns must hold zeros wherever the matrix is a K operand (the writers see to that).
// Coordinates are {0, row, col / 64}; box = 64 x box_rows x 1.
inline int make_tmap_bf16_slab(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  auto fn = tmap_encode_fn();
  if (!fn) return -1;
  if (reinterpret_cast<uintptr_t>(ptr) & 127) return -2;
  cuuint64_t gdim[3] = {64, rows, (cols + 63) / 64};
  cuuint64_t gstride[2] = {128, rows * 128};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -3;
}

// Row-major bf16 output [rows, cols] written in 32-row x 64-column slabs by the epilogue warps (128B swizzle).
inline int make_store_tmap_bf16(CUtensorMap* out, void* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
  return make_tmap_bf16_2d(out, ptr, rows, cols, ld, 32);
}
inline int make_store_tmap_bf16_slab(CUtensorMap* out, void* ptr, uint64_t rows, uint64_t cols) {
  return make_tmap_bf16_slab(out, ptr, rows, cols, 32);
}

// Row-major bf16 output written in 32-row x 32-column chunks (2 KB staging tiles, 64B swizzle); coordinates {col, row}.
inline int make_store_tmap_bf16_chunk(CUtensorMap* out, void* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
  auto fn = tmap_encode_fn();
  if (!fn) return -1;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15)) return -2;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -3;
}

// Slab-major store map for 32-column x 32-row chunks (2 KB staging tiles, 64B swizzle): coordinates {col % 64, row, col / 64}.
inline int make_store_tmap_bf16_slab32(CUtensorMap* out, void* ptr, uint64_t rows, uint64_t cols) {
  auto fn = tmap_encode_fn();
  if (!fn) return -1;
  if (reinterpret_cast<uintptr_t>(ptr) & 127) return -2;
  cuuint64_t gdim[3] = {64, rows, (cols + 63) / 64};
  cuuint64_t gstride[2] = {128, rows * 128};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -3;
}

// The caller's NCHW bf16 tensor [B, C, HW] as a 3-D map {HW, C, B}; box = 32 positions x 32 channels x 1 image, no
// swizzle (the epilogue stages channel-major tiles).  Needs HW % 8 == 0 (16-byte row pitch) and a 16-byte aligned base.
inline int make_tmap_nchw_bf16(CUtensorMap* out, void* ptr, uint64_t B, uint64_t C, uint64_t HW) {
  auto fn = tmap_encode_fn();
  if (!fn) return -1;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (HW % 8)) return -2;
  cuuint64_t gdim[3] = {HW, C, B};
  cuuint64_t gstride[2] = {HW * 2, C * HW * 2};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -3;
}

// Channel-major bf16 copy of the decoder output, [C][ld] with ld >= T a multiple of 8: the fused decoder epilogue
// stores its channel-major 32 x 32 tiles here when the caller's NCHW rows are not TMA-addressable (HW % 4 != 0, e.g.
// 7x7 maps); coordinates {token, channel}, box 32 x 32, no swizzle.
inline int make_store_tmap_bf16_cmajor(CUtensorMap* out, void* ptr, uint64_t C, uint64_t T, uint64_t ld) {
  auto fn = tmap_encode_fn();
  if (!fn) return -1;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 8)) return -2;
  cuuint64_t gdim[2] = {T, C};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -3;
}

constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = -1;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}
inline int device_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// A: K-major -> memory [M, K] pitch lda;  MN-major -> memory [K, M] pitch lda.   B likewise with N.
// k_splits_req <= 0 picks a split count that fills the machine when there are few output tiles.
// Returns 0 or a negative svb error code; *splits_out receives the number of split-K slices used.
template <int BLOCK_N, bool A_MN, bool B_MN, class Epi, bool BSTAT = false>
int launch_gemm(cudaStream_t stream, const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K,
                int k_splits_req, const typename Epi::Params& ep, int* splits_out = nullptr, int max_ctas = 0,
                unsigned long long a_policy = 0, bool a_slab = false, bool b_slab = false, int a_prefetch = 0,
                bool reverse_m = false) {
  using Cfg = GemmCfg<BLOCK_N, Epi::kSmemBytes, BSTAT>;
  if (M <= 0 || N <= 0 || K <= 0 || (N % 8)) return -2;  // pitches are validated by the tensor-map encoder
  CUtensorMap tmA, tmB;
  int rc;
  // slab-major operands: the logical matrix is [M, K] (K-major) or [K, M] (MN-major); lda / ldb are ignored
  if (a_slab) rc = !A_MN ? make_tmap_bf16_slab(&tmA, A, M, K, kBlockM) : make_tmap_bf16_slab(&tmA, A, K, M, kBlockK);
  else if (!A_MN) rc = make_tmap_bf16_2d(&tmA, A, M, K, lda, kBlockM);
  else rc = make_tmap_bf16_2d(&tmA, A, K, M, lda, kBlockK);
  if (rc) return rc;
  if (b_slab) rc = !B_MN ? make_tmap_bf16_slab(&tmB, B, N, K, BLOCK_N) : make_tmap_bf16_slab(&tmB, B, K, N, kBlockK);
  else if (!B_MN) rc = make_tmap_bf16_2d(&tmB, B, N, K, ldb, BLOCK_N);
  else rc = make_tmap_bf16_2d(&tmB, B, K, N, ldb, kBlockK);
  if (rc) return rc;

  GemmProblem p;
  p.M = M; p.N = N; p.K = K;
  p.a_policy = a_policy;
  p.a_slab = a_slab ? 1 : 0;
  p.b_slab = b_slab ? 1 : 0;
  p.a_prefetch = a_prefetch;
  p.reverse_m = reverse_m ? 1 : 0;
#ifdef SVB_GEMM_TRACE
  p.trace = gemm_trace_ptr();
#endif
  p.tiles_m = (M + kBlockM - 1) / kBlockM;
  p.tiles_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int kblocks = (K + kBlockK - 1) / kBlockK;
  const int sms = max_ctas > 0 ? max_ctas : device_sm_count();
  int splits = k_splits_req;
  if (splits <= 0) {
    const int mn_tiles = p.tiles_m * p.tiles_n;
    splits = mn_tiles >= sms ? 1 : (sms / mn_tiles);
  }
  if (splits > kblocks) splits = kblocks;
  if (splits < 1) splits = 1;
  const int kb_per = (kblocks + splits - 1) / splits;
  splits = (kblocks + kb_per - 1) / kb_per;  // drop empty slices
  p.k_splits = splits;
  p.k_per_split = kb_per * kBlockK;
  if (splits_out) *splits_out = splits;

  auto kern = gemm_bf16_kernel<BLOCK_N, A_MN, B_MN, Epi, BSTAT>;
  const uint32_t smem = Cfg::kSmemBytes;
  // the attribute is per device: a process that drives several GPUs must set it on each of them
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return -4;
  if (!configured[dev]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -4;
    configured[dev] = true;
  }
  const int num_tiles = p.tiles_m * p.tiles_n * p.k_splits;
  int grid = num_tiles < sms ? num_tiles : sms;
  if (BSTAT) {  // whole groups of tiles_n CTAs, each group walking a strided set of M tiles
    if (p.k_splits != 1 || K > Cfg::kResidentKBlocks * kBlockK || p.tiles_n > sms) return -5;
    int groups = sms / p.tiles_n;
    if (groups > p.tiles_m) groups = p.tiles_m;
    grid = groups * p.tiles_n;
  }
  (kern<<<grid, 64 + Epi::kWarps * 32, smem, stream>>>(tmA, tmB, p, ep), svb::count_launch());
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

// Split count the launcher would choose (needed to size split-K workspaces before launching).
template <int BLOCK_N>
inline int planned_splits(int M, int N, int K, int k_splits_req, int max_ctas = 0) {
  const int tiles_m = (M + kBlockM - 1) / kBlockM, tiles_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int kblocks = (K + kBlockK - 1) / kBlockK;
  const int sms = max_ctas > 0 ? max_ctas : device_sm_count();
  int splits = k_splits_req;
  if (splits <= 0) {
    const int mn_tiles = tiles_m * tiles_n;
    splits = mn_tiles >= sms ? 1 : (sms / mn_tiles);
  }
  if (splits > kblocks) splits = kblocks;
  if (splits < 1) splits = 1;
  const int kb_per = (kblocks + splits - 1) / splits;
  return (kblocks + kb_per - 1) / kb_per;
}

}  // namespace svb
