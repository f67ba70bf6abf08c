// Stand-alone bring-up test for the tcgen05 GEMM core (not part of the shipped library).
//   gemm_selftest <case>      exact integer-valued check against a CPU triple loop
//   gemm_selftest perf        times the encoder-shaped GEMM at cfg2 size
// Inputs are small integers so fp32 accumulation is exact and any layout/descriptor bug shows as a mismatch.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include <string>
#include "../sparse_vision_b200/csrc/gemm_host.cuh"
#include "../sparse_vision_b200/csrc/epilogues.cuh"

using namespace svb;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)(u >> 16);  // exact for small integers
}
static uint32_t rng_state = 12345u;
static int rnd_int() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return (int)((rng_state >> 24) % 7) - 3;
}

struct Case {
  const char* name;
  int M, N, K;
  bool a_mn, b_mn;
  int bn;
  int splits;  // 0 = auto, 1 = none
  bool bf16_out = false;  // bf16 output through the TMA slab store (K must keep |sums| <= 256 so bf16 is exact)
  bool bstat = false;     // B-stationary schedule (K <= 256)
};

template <int BN, bool AMN, bool BMN>
static int run(const Case& c) {
  // logical A[M,K], B[N,K]; memory layout depends on major-ness
  std::vector<float> A((size_t)c.M * c.K), B((size_t)c.N * c.K);
  for (auto& v : A) v = (float)rnd_int();
  for (auto& v : B) v = (float)rnd_int();
  std::vector<uint16_t> Ah(A.size()), Bh(B.size());
  for (int m = 0; m < c.M; ++m)
    for (int k = 0; k < c.K; ++k) Ah[AMN ? (size_t)k * c.M + m : (size_t)m * c.K + k] = f2bf(A[(size_t)m * c.K + k]);
  for (int n = 0; n < c.N; ++n)
    for (int k = 0; k < c.K; ++k) Bh[BMN ? (size_t)k * c.N + n : (size_t)n * c.K + k] = f2bf(B[(size_t)n * c.K + k]);
  void *dA, *dB;
  float* dC;
  CK(cudaMalloc(&dA, Ah.size() * 2));
  CK(cudaMalloc(&dB, Bh.size() * 2));
  CK(cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice));
  const int splits = planned_splits<BN>(c.M, c.N, c.K, c.splits);
  CK(cudaMalloc(&dC, (size_t)splits * c.M * c.N * 4));
  CK(cudaMemset(dC, 0xFF, (size_t)splits * c.M * c.N * 4));
  int used = 0;
  int rc;
  if (c.bf16_out) {
   if constexpr (BN == 256) {
    EpiStore::Params ep;
    memset(&ep, 0, sizeof(ep));
    ep.out = dC; ep.ld = c.N; ep.alpha = 1.0f; ep.out_bf16 = 1;
    if (make_store_tmap_bf16(&ep.tm, dC, c.M, c.N, c.N) == 0) ep.tm_valid = 1;
    else { printf("[%s] store tensor map failed\n", c.name); return 1; }
    if (c.bstat) rc = launch_gemm<BN, AMN, BMN, EpiStore, true>(0, dA, AMN ? c.M : c.K, dB, BMN ? c.N : c.K, c.M, c.N, c.K, 1, ep, &used);
    else rc = launch_gemm<BN, AMN, BMN, EpiStore>(0, dA, AMN ? c.M : c.K, dB, BMN ? c.N : c.K, c.M, c.N, c.K, 1, ep, &used);
   } else { printf("[%s] bf16 slab output needs BLOCK_N=256\n", c.name); return 1; }
  } else {
    EpiPartial::Params ep{dC, c.N, (long long)c.M * c.N};
    rc = launch_gemm<BN, AMN, BMN, EpiPartial>(0, dA, AMN ? c.M : c.K, dB, BMN ? c.N : c.K, c.M, c.N, c.K, c.splits, ep,
                                               &used);
  }
  if (rc) {
    printf("[%s] launch failed rc=%d\n", c.name, rc);
    return 1;
  }
  CK(cudaDeviceSynchronize());
  std::vector<float> C((size_t)used * c.M * c.N);
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  double maxerr = 0;
  for (int m = 0; m < c.M; ++m)
    for (int n = 0; n < c.N; ++n) {
      double ref = 0;
      for (int k = 0; k < c.K; ++k) ref += (double)A[(size_t)m * c.K + k] * B[(size_t)n * c.K + k];
      double got = 0;
      if (c.bf16_out) {
        const uint16_t* hb = reinterpret_cast<const uint16_t*>(C.data());
        uint32_t u = (uint32_t)hb[(size_t)m * c.N + n] << 16;
        float f;
        memcpy(&f, &u, 4);
        got = f;
      } else
        for (int s = 0; s < used; ++s) got += C[((size_t)s * c.M + m) * c.N + n];
      double err = fabs(got - ref);
      if (!(err <= 1e-3)) {
        if (bad < 8) printf("  mismatch m=%d n=%d got=%f ref=%f\n", m, n, got, ref);
        ++bad;
      }
      if (err > maxerr) maxerr = err;
    }
  printf("[%s] M=%d N=%d K=%d aMN=%d bMN=%d BN=%d splits=%d : %s (bad=%lld maxerr=%g)\n", c.name, c.M, c.N, c.K, AMN,
         BMN, BN, used, bad ? "FAIL" : "PASS", bad, maxerr);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return bad ? 1 : 0;
}

static int dispatch(const Case& c) {
  if (c.bn == 256) {
    if (!c.a_mn && !c.b_mn) return run<256, false, false>(c);
    if (!c.a_mn && c.b_mn) return run<256, false, true>(c);
    if (c.a_mn && c.b_mn) return run<256, true, true>(c);
    return run<256, true, false>(c);
  } else {
    if (!c.a_mn && !c.b_mn) return run<128, false, false>(c);
    if (!c.a_mn && c.b_mn) return run<128, false, true>(c);
    if (c.a_mn && c.b_mn) return run<128, true, true>(c);
    return run<128, true, false>(c);
  }
}

static const Case kCases[] = {
    {"kk_one_tile", 128, 256, 64, false, false, 256, 1},
    {"kk_mtail", 300, 256, 256, false, false, 256, 1},
    {"kk_deepk", 256, 512, 2048, false, false, 256, 1},
    {"kk_bn128", 300, 384, 192, false, false, 128, 1},
    {"kk_ntail", 200, 200, 72, false, false, 256, 1},
    {"kk_tiny_cfg1", 64, 64, 16, false, false, 128, 1},
    {"kk_many_tiles", 128 * 400 + 5, 256, 64, false, false, 256, 1},
    {"kmn_dE", 300, 512, 256, false, true, 256, 1},
    {"kmn_bn128", 130, 136, 80, false, true, 128, 1},
    {"mnmn_splitk", 256, 512, 3000, true, true, 256, 3},
    {"mnmn_auto", 256, 2048, 6272, true, true, 256, 0},
    {"mnmn_tail", 200, 264, 1000, true, true, 256, 0},
    {"mnk", 136, 128, 520, true, false, 128, 2},
    {"kk_bf16_tma", 300, 256, 16, false, false, 256, 1, true},
    {"kk_bf16_tma_ntail", 200, 200, 24, false, false, 256, 1, true},
    {"kk_bf16_tma_n96", 130, 96, 16, false, false, 256, 1, true},
    {"kk_bf16_tma_many", 128 * 300 + 40, 512, 16, false, false, 256, 1, true},
    {"kk_bstat", 128 * 300 + 40, 512, 16, false, false, 256, 1, true, true},
    {"kk_bstat_k24_ntail", 5000, 200, 24, false, false, 256, 1, true, true},
    {"kmn_bstat", 3000, 768, 16, false, true, 256, 1, true, true},
};

static int perf() {
  // encoder-shaped GEMM at cfg2: T=200704 tokens, C=256, F=2048, bias+relu, bf16 out
  const int M = 200704, N = 2048, K = 256;
  void *dA, *dB, *dE;
  float* dbias;
  CK(cudaMalloc(&dA, (size_t)M * K * 2));
  CK(cudaMalloc(&dB, (size_t)N * K * 2));
  CK(cudaMalloc(&dE, (size_t)M * N * 2));
  CK(cudaMalloc(&dbias, N * 4));
  CK(cudaMemset(dA, 0x3c, (size_t)M * K * 2));
  CK(cudaMemset(dB, 0x3c, (size_t)N * K * 2));
  CK(cudaMemset(dbias, 0, N * 4));
  EpiStore::Params ep;
  memset(&ep, 0, sizeof(ep));
  ep.out = dE; ep.ld = N; ep.bias = dbias; ep.alpha = 1.0f; ep.relu = 1; ep.out_bf16 = 1;
  if (make_store_tmap_bf16(&ep.tm, dE, M, N, N) == 0) ep.tm_valid = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) launch_gemm<256, false, false, EpiStore>(0, dA, K, dB, K, M, N, K, 1, ep);
  CK(cudaDeviceSynchronize());
  const int iters = 10;
  cudaEventRecord(e0);
  for (int it = 0; it < iters; ++it) launch_gemm<256, false, false, EpiStore>(0, dA, K, dB, K, M, N, K, 1, ep);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= iters;
  printf("[perf enc] %.3f ms  %.1f TFLOP/s  out %.1f GB/s\n", ms, 2.0 * M * N * K / ms * 1e-9,
         (double)M * N * 2 / ms * 1e-6);
  for (int it = 0; it < 3; ++it) launch_gemm<256, false, false, EpiStore, true>(0, dA, K, dB, K, M, N, K, 1, ep);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int it = 0; it < iters; ++it) launch_gemm<256, false, false, EpiStore, true>(0, dA, K, dB, K, M, N, K, 1, ep);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= iters;
  printf("[perf enc B-stationary] %.3f ms  %.1f TFLOP/s  out %.1f GB/s\n", ms, 2.0 * M * N * K / ms * 1e-9,
         (double)M * N * 2 / ms * 1e-6);
  // split-K weight-gradient shape: dW_dec[C,F] = diff^T[C,T] e[T,F]
  {
    const int M2 = 256, N2 = 2048, K2 = 200704;
    float* dP;
    const int splits = planned_splits<256>(M2, N2, K2, 0);
    CK(cudaMalloc(&dP, (size_t)splits * M2 * N2 * 4));
    EpiPartial::Params ep2{dP, N2, (long long)M2 * N2};
    for (int it = 0; it < 3; ++it) launch_gemm<256, true, true, EpiPartial>(0, dA, M2, dE, N2, M2, N2, K2, 0, ep2);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int it = 0; it < iters; ++it) launch_gemm<256, true, true, EpiPartial>(0, dA, M2, dE, N2, M2, N2, K2, 0, ep2);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    printf("[perf dWdec splits=%d] %.3f ms  %.1f TFLOP/s\n", splits, ms, 2.0 * M2 * N2 * K2 / ms * 1e-9);
  }
  return 0;
}

// EpiEnc feature toggles: which fused output costs what (cfg2 shape)
static int perf_enc_variants() {
  const int M = 200704, N = 2048, K = 256, HW = 784, words = N / 32;
  void *dA, *dB, *dE;
  float *dbias, *dl1;
  uint32_t *dact, *dmask;
  CK(cudaMalloc(&dA, (size_t)M * K * 2));
  CK(cudaMalloc(&dB, (size_t)N * K * 2));
  CK(cudaMalloc(&dE, (size_t)M * N * 2));
  CK(cudaMalloc(&dbias, N * 4));
  CK(cudaMalloc(&dl1, (size_t)(M / 128 + 1) * (N / 256) * 16 * 4));
  CK(cudaMalloc(&dact, (size_t)(M / HW) * words * 4));
  CK(cudaMalloc(&dmask, (size_t)M * words * 4));
  CK(cudaMemset(dA, 0x3c, (size_t)M * K * 2));
  CK(cudaMemset(dB, 0x3c, (size_t)N * K * 2));
  CK(cudaMemset(dbias, 0, N * 4));
  CK(cudaMemset(dact, 0, (size_t)(M / HW) * words * 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int variant = 0; variant < 5; ++variant) {
    EpiEnc::Params ep;
    memset(&ep, 0, sizeof(ep));
    ep.bias = dbias; ep.e_bf16 = (__nv_bfloat16*)dE; ep.hw = HW; ep.words = words;
    make_store_tmap_bf16(&ep.tm_e, dE, M, N, N);
    if (variant == 1 || variant == 4) ep.act_bits = dact;
    if (variant == 2 || variant == 4) ep.mask_words = dmask;
    if (variant == 3 || variant == 4) ep.l1_partial = dl1;
    for (int it = 0; it < 3; ++it) launch_gemm<256, false, false, EpiEnc, true>(0, dA, K, dB, K, M, N, K, 1, ep);
    CK(cudaDeviceSynchronize());
    const int iters = 10;
    cudaEventRecord(e0);
    for (int it = 0; it < iters; ++it) launch_gemm<256, false, false, EpiEnc, true>(0, dA, K, dB, K, M, N, K, 1, ep);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    printf("[perf EpiEnc variant %d: act=%d mask=%d l1=%d] %.3f ms  %.1f TFLOP/s\n", variant, ep.act_bits != nullptr,
           ep.mask_words != nullptr, ep.l1_partial != nullptr, ms, 2.0 * M * N * K / ms * 1e-9);
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    printf("%d\n", (int)(sizeof(kCases) / sizeof(kCases[0])));
    return 0;
  }
  if (std::string(argv[1]) == "perf") return perf();
  if (std::string(argv[1]) == "perf_enc") return perf_enc_variants();
  const int i = atoi(argv[1]);
  if (i < 0 || i >= (int)(sizeof(kCases) / sizeof(kCases[0]))) return 3;
  return dispatch(kCases[i]);
}
