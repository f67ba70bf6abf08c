// Stand-alone check + timing of the fused encoder backward (sparse_vision_b200/csrc/fused_bwd_sm100.cuh) against naive
// kernels on integer-valued inputs (every sum is an integer < 2^24, so fp32 accumulation is exact in any order and the
// comparison is bit-exact).
//   fused_bwd_probe check     several small shapes (row-major and slab-major operands, T / F tails, C = 64..256)
//   fused_bwd_probe perf      cfg2 size (T = 200704, C = 256, F = 2048)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define SVB_FBW_TRACE 1
#include "../sparse_vision_b200/csrc/fused_bwd_sm100.cuh"

using namespace svb;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

namespace {

// DP[t][f] = mask ? sum_c DIFF[t,c] W[c,f] + l1c : 0 (rounded to bf16 like the P tile), fp32 copy for the column sums
__global__ void ref_dpre_kernel(const __nv_bfloat16* DIFF, const __nv_bfloat16* W, const uint32_t* mask, float l1c,
                                __nv_bfloat16* DP, float* DPf, int T, int C, int F) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (f >= F || t >= T) return;
  float a = 0.f;
  for (int c = 0; c < C; ++c) a += __bfloat162float(DIFF[(size_t)t * C + c]) * __bfloat162float(W[(size_t)c * F + f]);
  const uint32_t word = mask[mask_index(t, f >> 5, T)];
  const float v = (word >> (f & 31)) & 1u ? a + l1c : 0.f;
  DP[(size_t)t * F + f] = __float2bfloat16_rn(v);
  DPf[(size_t)t * F + f] = v;
}
__global__ void ref_dw_kernel(const __nv_bfloat16* DP, const float* DPf, const __nv_bfloat16* X, float* dW, float* cs,
                              int T, int C, int F) {
  const int c = threadIdx.x, f = blockIdx.x;
  float a = 0.f, s = 0.f;
  for (int t = 0; t < T; ++t) {
    a += __bfloat162float(DP[(size_t)t * F + f]) * __bfloat162float(X[(size_t)t * C + c]);
    s += DPf[(size_t)t * F + f];
  }
  dW[(size_t)f * C + c] = a;
  if (c == 0) cs[f] = s;
}
// row-major [T, C] -> slab-major [C/64][T][64]
std::vector<__nv_bfloat16> to_slab(const std::vector<__nv_bfloat16>& a, int T, int C) {
  std::vector<__nv_bfloat16> o((size_t)T * C);
  for (int t = 0; t < T; ++t)
    for (int c = 0; c < C; ++c) o[slab_offset(t, c, T)] = a[(size_t)t * C + c];
  return o;
}

uint32_t rng = 2463534242u;
uint32_t rnd() { rng ^= rng << 13; rng ^= rng >> 17; rng ^= rng << 5; return rng; }
int rnd3() { return (int)(rnd() % 3) - 1; }

int run(int T, int C, int F, bool slab, bool check, int iters) {
  std::vector<__nv_bfloat16> hD((size_t)T * C), hX((size_t)T * C), hW((size_t)C * F);
  for (auto& v : hD) v = __float2bfloat16((float)rnd3());
  for (auto& v : hX) v = __float2bfloat16((float)rnd3());
  for (auto& v : hW) v = __float2bfloat16((float)rnd3());
  const int words = (F + 31) / 32;
  std::vector<uint32_t> hM((size_t)T * 4 * ((words + 3) / 4), 0u);
  for (int t = 0; t < T; ++t)
    for (int w = 0; w < words; ++w) hM[mask_index(t, w, T)] = rnd() & rnd();   // ~25 % active
  const int slots = fused_bwd_slots(T, C, F);
  if (!fused_bwd_supported(T, C, F)) { printf("shape not supported\n"); return 2; }
  __nv_bfloat16 *dD, *dX, *dW, *dDr, *dXr;
  uint32_t* dM;
  float *dPart, *dCs;
  CK(cudaMalloc(&dD, hD.size() * 2)); CK(cudaMalloc(&dX, hX.size() * 2)); CK(cudaMalloc(&dW, hW.size() * 2));
  CK(cudaMalloc(&dDr, hD.size() * 2)); CK(cudaMalloc(&dXr, hX.size() * 2));
  CK(cudaMalloc(&dM, hM.size() * 4));
  // both outputs get a 4 KB guard band behind them (0xA5 pattern) that the kernel must leave untouched
  const size_t nPart = (size_t)slots * F * C, nCs = (size_t)2 * slots * F, kGuard = 1024;
  CK(cudaMalloc(&dPart, (nPart + kGuard) * 4)); CK(cudaMalloc(&dCs, (nCs + kGuard) * 4));
  CK(cudaMemset(dPart + nPart, 0xA5, kGuard * 4)); CK(cudaMemset(dCs + nCs, 0xA5, kGuard * 4));
  const std::vector<__nv_bfloat16> sD = slab ? to_slab(hD, T, C) : hD, sX = slab ? to_slab(hX, T, C) : hX;
  CK(cudaMemcpy(dD, sD.data(), sD.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dX, sX.data(), sX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dDr, hD.data(), hD.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dXr, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dM, hM.data(), hM.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dPart, 0xFF, (size_t)slots * F * C * 4)); CK(cudaMemset(dCs, 0xFF, (size_t)2 * slots * F * 4));
  const float l1c = 1.f;
  int rc = launch_fused_bwd(0, dW, dD, slab, C, dX, slab, C, dM, T, C, F, l1c, dPart, dCs);
  if (rc) { printf("launch failed: %d\n", rc); return 2; }
  CK(cudaDeviceSynchronize());
  if (check) {
    __nv_bfloat16* rDP; float *rDPf, *rdW, *rcs;
    CK(cudaMalloc(&rDP, (size_t)T * F * 2)); CK(cudaMalloc(&rDPf, (size_t)T * F * 4));
    CK(cudaMalloc(&rdW, (size_t)F * C * 4)); CK(cudaMalloc(&rcs, (size_t)F * 4));
    ref_dpre_kernel<<<dim3((F + 255) / 256, T), 256>>>(dDr, dW, dM, l1c, rDP, rDPf, T, C, F);
    ref_dw_kernel<<<F, C>>>(rDP, rDPf, dXr, rdW, rcs, T, C, F);
    CK(cudaDeviceSynchronize());
    std::vector<float> part((size_t)slots * F * C), cs((size_t)2 * slots * F), rw((size_t)F * C), rc_((size_t)F);
    CK(cudaMemcpy(part.data(), dPart, part.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cs.data(), dCs, cs.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rw.data(), rdW, rw.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rc_.data(), rcs, rc_.size() * 4, cudaMemcpyDeviceToHost));
    size_t bw = 0, bc = 0;
    {
      std::vector<uint32_t> g1(kGuard), g2(kGuard);
      CK(cudaMemcpy(g1.data(), dPart + nPart, kGuard * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(g2.data(), dCs + nCs, kGuard * 4, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < kGuard; ++i)
        if (g1[i] != 0xA5A5A5A5u || g2[i] != 0xA5A5A5A5u) { printf("guard band overwritten at word %zu\n", i); ++bw; break; }
    }
    for (size_t i = 0; i < rw.size(); ++i) {
      float s = 0.f;
      for (int k = 0; k < slots; ++k) s += part[(size_t)k * F * C + i];
      if (s != rw[i] && ++bw <= 5) printf("dW mismatch at (f=%zu,c=%zu): %g vs %g\n", i / C, i % C, s, rw[i]);
    }
    for (int f = 0; f < F; ++f) {
      float s = 0.f;
      for (int k = 0; k < 2 * slots; ++k) s += cs[(size_t)k * F + f];
      if (s != rc_[f] && ++bc <= 5) printf("colsum mismatch at f=%d: %g vs %g\n", f, s, rc_[f]);
    }
    printf("check T=%d C=%d F=%d %s %s slots=%d: dW mismatches %zu / %zu, colsum mismatches %zu / %d -> %s\n", T, C, F,
           slab ? "slab" : "rowmajor", fused_bwd_two_cta(C) ? "2-CTA" : "1-CTA", slots, bw, rw.size(), bc, F, (bw | bc) ? "FAIL" : "PASS");
    cudaFree(rDP); cudaFree(rDPf); cudaFree(rdW); cudaFree(rcs);
    rc = (bw | bc) ? 1 : 0;
  } else {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch_fused_bwd(0, dW, dD, slab, C, dX, slab, C, dM, T, C, F, l1c, dPart, dCs);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    {   // one more launch with the wait-cycle trace on: who waits for whom
      const int grid = (fused_bwd_two_cta(C) ? 2 * ((F + 255) / 256) : (F + 127) / 128) * slots;
      long long* dT;
      CK(cudaMalloc(&dT, (size_t)grid * 8 * 8)); CK(cudaMemset(dT, 0, (size_t)grid * 8 * 8));
      fused_bwd_trace_ptr() = dT;
      CK(cudaEventRecord(e0));
      launch_fused_bwd(0, dW, dD, slab, C, dX, slab, C, dM, T, C, F, l1c, dPart, dCs);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      fused_bwd_trace_ptr() = nullptr;
      float ms1 = 0; CK(cudaEventElapsedTime(&ms1, e0, e1));
      std::vector<long long> tr((size_t)grid * 8);
      CK(cudaMemcpy(tr.data(), dT, tr.size() * 8, cudaMemcpyDeviceToHost));
      double avg[8] = {0}; int cnt[8] = {0};
      for (int b = 0; b < grid; ++b) for (int q = 0; q < 8; ++q) if (tr[(size_t)b * 8 + q] || q < 2 || q >= 6) { avg[q] += tr[(size_t)b * 8 + q]; cnt[q]++; }
      const char* nm[8] = {"prod:d_empty", "prod:x_empty", "mma:acc1_empty", "mma:d_full", "mma:p_full", "mma:x_full", "epi:acc1_full", "epi:p_empty"};
      printf("  trace (traced launch %.4f ms = %.0f kcycles at 1.965 GHz); mean wait kcycles per CTA:", ms1, ms1 * 1965.0);
      for (int q = 0; q < 8; ++q) printf(" %s=%.0f", nm[q], cnt[q] ? avg[q] / cnt[q] / 1e3 : 0.0);
      printf("\n");
      cudaFree(dT);
    }
    printf("perf T=%d C=%d F=%d %s %s: %.4f ms per call = %.0f TFLOP/s over both GEMMs (un-fused dE + dW_enc in the step: ~0.40 ms)\n",
           T, C, F, slab ? "slab" : "rowmajor", fused_bwd_two_cta(C) ? "2-CTA" : "1-CTA", ms, 4.0 * T * C * F / (ms * 1e-3) * 1e-12);
    rc = 0;
  }
  cudaFree(dD); cudaFree(dX); cudaFree(dW); cudaFree(dDr); cudaFree(dXr); cudaFree(dM); cudaFree(dPart); cudaFree(dCs);
  return rc;
}

}  // namespace

int main(int argc, char** argv) {
  const char* mode = argc > 1 ? argv[1] : "check";
  if (!strcmp(mode, "check")) {
    int rc = run(384, 256, 768, false, true, 0);
    rc |= run(384, 256, 768, true, true, 0);
    rc |= run(128 * 40 + 50, 256, 2048, true, true, 0);    // T tail; several blocks per CTA: every barrier phase wraps
    rc |= run(128 * 37, 128, 1000, false, true, 0);        // F tail (1000 = 7 * 128 + 104), C = 128
    rc |= run(900, 64, 256, true, true, 0);
    rc |= run(128 * 21 + 7, 192, 1536, true, true, 0);
    return rc;
  }
  int rc = run(200704, 256, 2048, true, false, 20);
  rc |= run(200704, 256, 2048, false, false, 20);
  return rc;
}
