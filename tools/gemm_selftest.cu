// Stand-alone bring-up test for the tcgen05 GEMM core (not part of the shipped library).
//   gemm_selftest <case>      exact integer-valued check against a CPU triple loop
//   gemm_selftest perf        times the encoder-shaped GEMM at cfg2 size
// Inputs are small integers so fp32 accumulation is exact and any layout/descriptor bug shows as a mismatch.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include <string>
#include <type_traits>
#define SVB_GEMM_TRACE 1
#include "../sparse_vision_b200/csrc/gemm_host.cuh"
#include "../sparse_vision_b200/csrc/epilogues.cuh"
#include "../sparse_vision_b200/csrc/gemm2_sm100.cuh"

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
  bool two_cta = false;   // cta_group::2 B-stationary kernel (gemm2_sm100.cuh)
  bool ones = false;      // EpiPartialOnes: also check the row sums of A from the extra ones column
  bool two_stream = false;  // cta_group::2 STREAMING kernel (gemm2_stream_kernel): any K, split-K, either major-ness
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
  float* dRow = nullptr;
  CK(cudaMalloc(&dA, Ah.size() * 2));
  CK(cudaMalloc(&dB, Bh.size() * 2));
  CK(cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice));
  const int splits = c.two_stream ? planned_splits2(c.M, c.N, c.K, c.splits) : planned_splits<BN>(c.M, c.N, c.K, c.splits);
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
    if (c.two_stream) rc = launch_gemm2_stream<AMN, BMN, EpiStore>(0, dA, AMN ? c.M : c.K, dB, BMN ? c.N : c.K, c.M, c.N, c.K, 1, ep, &used);
    else if (c.two_cta) {
      if constexpr (!AMN) { rc = launch_gemm2_bstat<BMN, EpiStore>(0, dA, c.K, dB, BMN ? c.N : c.K, c.M, c.N, c.K, ep); used = 1; }
      else { printf("[%s] the 2-CTA kernel needs a K-major A\n", c.name); return 1; }
    } else if (c.bstat) rc = launch_gemm<BN, AMN, BMN, EpiStore, true>(0, dA, AMN ? c.M : c.K, dB, BMN ? c.N : c.K, c.M, c.N, c.K, 1, ep, &used);
    else rc = launch_gemm<BN, AMN, BMN, EpiStore>(0, dA, AMN ? c.M : c.K, dB, BMN ? c.N : c.K, c.M, c.N, c.K, 1, ep, &used);
   } else { printf("[%s] bf16 slab output needs BLOCK_N=256\n", c.name); return 1; }
  } else if (c.ones) {
   if constexpr (BN == 256) {
    CK(cudaMalloc(&dRow, (size_t)splits * c.M * 4));
    CK(cudaMemset(dRow, 0xFF, (size_t)splits * c.M * 4));
    EpiPartialOnes::Params ep{dC, c.N, (long long)c.M * c.N, dRow};
    rc = launch_gemm<BN, AMN, BMN, EpiPartialOnes>(0, dA, AMN ? c.M : c.K, dB, BMN ? c.N : c.K, c.M, c.N, c.K, c.splits, ep,
                                                   &used);
   } else { printf("[%s] the ones column needs BLOCK_N=256\n", c.name); return 1; }
  } else {
    EpiPartial::Params ep{dC, c.N, (long long)c.M * c.N};
    if (c.two_stream) {
      if constexpr (BN == 256) rc = launch_gemm2_stream<AMN, BMN, EpiPartial>(0, dA, AMN ? c.M : c.K, dB, BMN ? c.N : c.K, c.M, c.N, c.K, c.splits, ep, &used);
      else { printf("[%s] the 2-CTA kernels need BLOCK_N=256\n", c.name); return 1; }
    } else
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
  if (dRow) {
    std::vector<float> R((size_t)used * c.M);
    CK(cudaMemcpy(R.data(), dRow, R.size() * 4, cudaMemcpyDeviceToHost));
    for (int m = 0; m < c.M; ++m) {
      double ref = 0, got = 0;
      for (int k = 0; k < c.K; ++k) ref += A[(size_t)m * c.K + k];
      for (int s = 0; s < used; ++s) got += R[(size_t)s * c.M + m];
      if (!(fabs(got - ref) <= 1e-3)) {
        if (bad < 8) printf("  row-sum mismatch m=%d got=%f ref=%f\n", m, got, ref);
        ++bad;
      }
    }
    cudaFree(dRow);
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
    {"kk_2cta_one_pair", 256, 256, 16, false, false, 256, 1, true, true, true},
    {"kk_2cta", 128 * 300 + 40, 512, 16, false, false, 256, 1, true, true, true},
    {"kk_2cta_k24_ntail", 5000, 200, 24, false, false, 256, 1, true, true, true},
    {"kmn_2cta", 3000, 768, 16, false, true, 256, 1, true, true, true},
    {"kk_2cta_mtail_odd", 128 * 7 + 3, 256, 16, false, false, 256, 1, true, true, true},
    {"s2_kk_deepk", 256, 512, 2048, false, false, 256, 1, false, false, false, false, true},
    {"s2_kk_mtail", 300, 256, 256, false, false, 256, 1, false, false, false, false, true},
    {"s2_kk_many_tiles", 128 * 400 + 5, 256, 192, false, false, 256, 1, false, false, false, false, true},
    {"s2_kk_ntail", 700, 200, 72, false, false, 256, 1, false, false, false, false, true},
    {"s2_kmn_dE", 300, 512, 256, false, true, 256, 1, false, false, false, false, true},
    {"s2_mnmn_splitk", 256, 512, 3000, true, true, 256, 3, false, false, false, false, true},
    {"s2_mnmn_auto", 256, 2048, 6272, true, true, 256, 0, false, false, false, false, true},
    {"s2_mnmn_tail", 200, 264, 1000, true, true, 256, 0, false, false, false, false, true},
    {"s2_mnmn_m512", 512, 1024, 5000, true, true, 256, 0, false, false, false, false, true},
    {"s2_kk_bf16_tma_many", 128 * 300 + 40, 512, 16, false, false, 256, 1, true, false, false, false, true},
    {"mnmn_ones_splitk", 256, 512, 3000, true, true, 256, 3, false, false, false, true},
    {"mnmn_ones_auto_tail", 200, 264, 1000, true, true, 256, 0, false, false, false, true},
    {"kk_ones", 300, 256, 192, false, false, 256, 1, false, false, false, true},
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
  for (int it = 0; it < 3; ++it) launch_gemm2_bstat<false, EpiStore>(0, dA, K, dB, K, M, N, K, ep);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int it = 0; it < iters; ++it) launch_gemm2_bstat<false, EpiStore>(0, dA, K, dB, K, M, N, K, ep);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= iters;
  printf("[perf enc 2-CTA B-stationary] %.3f ms  %.1f TFLOP/s  out %.1f GB/s\n", ms, 2.0 * M * N * K / ms * 1e-9,
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

template <class T> struct TypeTag { using type = T; };
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
  auto run = [&](auto tag, bool two, int variant, bool slab) {
    using E = typename decltype(tag)::type;
    typename E::Params ep;
    memset(&ep, 0, sizeof(ep));
    ep.bias = dbias; ep.e_bf16 = (__nv_bfloat16*)dE; ep.words = words; ep.e_slab = slab ? 1 : 0;
    if (slab) make_store_tmap_bf16_slab32(&ep.tm_e, dE, M, N); else make_store_tmap_bf16_chunk(&ep.tm_e, dE, M, N, N);
    if (variant == 2 || variant == 4) ep.mask_words = dmask;
    if (variant == 3 || variant == 4) ep.l1_partial = dl1;
    auto launch = [&]() {
      if (two) launch_gemm2_bstat<false, E>(0, dA, K, dB, K, M, N, K, ep);
      else launch_gemm<256, false, false, E, true>(0, dA, K, dB, K, M, N, K, 1, ep);
    };
    for (int it = 0; it < 3; ++it) launch();
    CK(cudaDeviceSynchronize());
    const int iters = 10;
    cudaEventRecord(e0);
    for (int it = 0; it < iters; ++it) launch();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    printf("[perf %s %s variant %d: mask=%d l1=%d %s] %.3f ms  %.1f TFLOP/s\n", two ? "2-CTA" : "1-CTA",
           sizeof(typename E::Writer) > sizeof(ChunkWriter) ? "EpiEnc4" : "EpiEnc", variant, ep.mask_words != nullptr,
           ep.l1_partial != nullptr, slab ? "slab" : "rowmajor", ms, 2.0 * M * N * K / ms * 1e-9);
    long long* tr;
    CK(cudaMalloc(&tr, 148 * 4 * sizeof(long long)));
    CK(cudaMemset(tr, 0, 148 * 4 * sizeof(long long)));
    gemm_trace_ptr() = tr;
    launch();
    CK(cudaDeviceSynchronize());
    gemm_trace_ptr() = nullptr;
    long long h[148 * 4];
    CK(cudaMemcpy(h, tr, sizeof(h), cudaMemcpyDeviceToHost));
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int n = 0, nl = 0;
    for (int b = 0; b < 144; ++b) {
      s0 += h[b * 4]; s3 += h[b * 4 + 3]; ++n;
      if (!two || (b & 1) == 0) { s1 += h[b * 4 + 1]; s2 += h[b * 4 + 2]; ++nl; }
    }
    printf("    wait kcycles per CTA: producer on free slot %.0f | MMA on operands %.0f | MMA on free accumulator %.0f | epilogue warp 0 on accumulator %.0f   (kernel ~%.0f kcycles)\n",
           s0 / n / 1e3, s1 / nl / 1e3, s2 / nl / 1e3, s3 / n / 1e3, ms * 1.965e6 / 1e3);
    cudaFree(tr);
  };
  for (int variant : {0, 2, 3, 4}) {
    run(TypeTag<EpiEnc>{}, false, variant, true);
    run(TypeTag<EpiEnc>{}, true, variant, true);
    run(TypeTag<EpiEnc4>{}, true, variant, true);
  }
  return 0;
}

// Bring-up probe: what bounds a K=256 tile?  MODE 0 = epilogue does nothing (no TMEM read), 1 = TMEM read only,
// 2 = TMEM read + bf16 pack + TMA slab store, 3 = TMEM read + ALU work (relu, mask bits, sum) without stores.
template <int MODE, int WARPS>
struct EpiProbe {
  struct Params {
    alignas(64) CUtensorMap tm;
    float* sink;
  };
  static constexpr int kWarps = WARPS;
  static constexpr int kColVecs = 0;
  static constexpr bool kSkipAccLoad = (MODE == 0);
  static constexpr uint32_t kSmemBytes = MODE == 16 ? WARPS * 2048 : SlabWriter1::bytes(WARPS);  // 16: per-chunk tiles -> 5 A stages
  const Params& p;
  SlabWriter1 slab;
  float sum;
  uint32_t bits;
  int ew;
  __device__ EpiProbe(const Params& p_, uint8_t* smem, int ew_, int) : p(p_), sum(0.f), bits(0), ew(ew_) { slab.init(smem, ew_); }
  __device__ void colvec_fetch(const GemmProblem&, const TileInfo&, int) {}
  __device__ void colvec_commit(uint32_t, int) {}
  __device__ void begin_tile(const GemmProblem&, const TileInfo&, int, int, int) {}
  __device__ __forceinline__ void chunk(const GemmProblem& g, const TileInfo& ti, int row, int col0, float (&v)[32], int wq, int lane, int) {
    if (MODE == 2 || MODE == 9 || MODE == 10) {
      const int half = (col0 >> 5) & 1;
      const bool skip = MODE == 10 && ((col0 >> 6) & 1);
      if (!skip) {
        slab.put(half, lane, v);
        if (half == 1) slab.flush(&p.tm, col0 - 32, MODE == 9 ? ((ti.m0 & 4095) + wq * 32) : (ti.m0 + wq * 32), lane);
      }
    }
    if (MODE == 12 || MODE == 13) {  // row-major slabs with an evict-first hint on the stores (13: plus evict-last A loads)
      const int half = (col0 >> 5) & 1;
      if (half == 0) { if (lane == 0) bulk_wait_read<0>(); __syncwarp(); }
      uint8_t* rowp = slab.base + lane * 128;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = half * 4 + i;
        *reinterpret_cast<uint4*>(rowp + ((j ^ (lane & 7)) << 4)) = pack8_bf16(v + 8 * i);
      }
      if (half == 1) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { tma_store_2d_hint(&p.tm, slab.base, col0 - 32, ti.m0 + wq * 32, kL2EvictFirst); bulk_commit(); }
      }
    }
    if (MODE == 14 || MODE == 15) {  // no shared memory at all: packed bf16 straight from registers into the slab-major
      // layout, every lane writes the 64 bytes of its row with two 32-byte stores (14) or four 16-byte stores (15)
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      const int r = ti.m0 + wq * 32 + lane;
      if (r < g.M) {
        char* dst = reinterpret_cast<char*>(p.sink) +
                    ((static_cast<size_t>(col0 >> 6) * g.M + r) * 64 + (col0 & 63)) * 2;
        if (MODE == 14) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 32 * h), "r"(w[8 * h]),
                         "r"(w[8 * h + 1]), "r"(w[8 * h + 2]), "r"(w[8 * h + 3]), "r"(w[8 * h + 4]), "r"(w[8 * h + 5]),
                         "r"(w[8 * h + 6]), "r"(w[8 * h + 7])
                         : "memory");
        } else {
#pragma unroll
          for (int h = 0; h < 4; ++h)
            *reinterpret_cast<uint4*>(dst + 16 * h) = make_uint4(w[4 * h], w[4 * h + 1], w[4 * h + 2], w[4 * h + 3]);
        }
      }
    }
    if (MODE == 16) {  // slab-major output staged per 32-column chunk (2 KB per warp, 64B swizzle): one more A stage fits
      uint8_t* buf = slab.base - ew * 4096 + ew * 2048;
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      uint8_t* frow = buf + lane * 64;
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(frow + ((i ^ ((lane >> 1) & 3)) << 4)) = pack8_bf16(v + 8 * i);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { tma_store_3d(&p.tm, buf, col0 & 63, ti.m0 + wq * 32, col0 >> 6); bulk_commit(); }
    }
    if (MODE == 11) {  // slab-major output [N/64][M][64]: every 32 x 64 slab is 4 KB of contiguous memory
      const int half = (col0 >> 5) & 1;
      if (half == 0) { if (lane == 0) bulk_wait_read<0>(); __syncwarp(); }
      uint8_t* rowp = slab.base + lane * 128;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = half * 4 + i;
        *reinterpret_cast<uint4*>(rowp + ((j ^ (lane & 7)) << 4)) = pack8_bf16(v + 8 * i);
      }
      if (half == 1) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(&p.tm),
                       "r"(smem_u32(slab.base)), "r"(0), "r"(ti.m0 + wq * 32), "r"((col0 - 32) >> 6) : "memory");
          bulk_commit();
        }
      }
    }
    if (MODE >= 4 && MODE <= 6) {  // 4 = STS only, 5 = TMA store only (with the read wait), 6 = STS + TMA store without the read wait
      const int half = (col0 >> 5) & 1;
      if (MODE == 5 && half == 0) { if (lane == 0) bulk_wait_read<0>(); __syncwarp(); }
      if (MODE != 5) {
        uint8_t* rowp = slab.base + lane * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = half * 4 + i;
          *reinterpret_cast<uint4*>(rowp + ((j ^ (lane & 7)) << 4)) = pack8_bf16(v + 8 * i);
        }
      }
      if (MODE != 4 && half == 1) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) { tma_store_2d(&p.tm, slab.base, col0 - 32, ti.m0 + wq * 32); bulk_commit(); }
      }
    }
    if (MODE == 7 || MODE == 8) {  // STS into the slab, then coalesced st.global.v4 (4 rows x 128 B per instruction); 8 = small L2-resident window
      const int half = (col0 >> 5) & 1;
      uint8_t* rowp = slab.base + lane * 128;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = half * 4 + i;
        *reinterpret_cast<uint4*>(rowp + ((j ^ (lane & 7)) << 4)) = pack8_bf16(v + 8 * i);
      }
      if (half == 1) {
        __syncwarp();
        const int r0 = ti.m0 + wq * 32;
        const int rwrap = MODE == 8 ? (r0 & 4095) : r0;
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.sink) + 1024;  // sink doubles as the output base (offset keeps slot 0 free)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + (lane >> 3), j = lane & 7;
          const uint4 q = *reinterpret_cast<const uint4*>(slab.base + r * 128 + ((j ^ (r & 7)) << 4));
          *reinterpret_cast<uint4*>(out + (size_t)(rwrap + r) * g.N + (col0 - 32) + j * 8) = q;
        }
        __syncwarp();
      }
    }
    if (MODE == 3) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float t = 1.0f - v[j];
        bits = __funnelshift_l(__float_as_uint(t), bits, 1);
        sum += fmaxf(-t, 0.f);
      }
    }
  }
  __device__ void end_tile(const GemmProblem&, const TileInfo&, int, int, int) {}
  __device__ void finish(int, int lane) {
    if (MODE == 2 || (MODE >= 5 && MODE != 7 && MODE != 8 && MODE != 14 && MODE != 15)) slab.drain(lane);  // (modes 11, 16 included)
    if (MODE == 3 && sum == 12345.678f && bits == 77) p.sink[0] = sum;
  }
};

static int g_prefetch = 0;
template <int MODE, int WARPS>
static void probe_one(const void* dA, const void* dB, void* dE, float* sink, int M, int N, int K) {
  typename EpiProbe<MODE, WARPS>::Params ep;
  memset(&ep, 0, sizeof(ep));
  ep.sink = sink;
  make_store_tmap_bf16(&ep.tm, dE, M, N, N);
  if (MODE == 16) make_store_tmap_bf16_slab32(&ep.tm, dE, M, N);
  if (MODE == 11) {
    cuuint64_t gdim[3] = {64, (cuuint64_t)M, (cuuint64_t)N / 64};
    cuuint64_t gstride[2] = {128, (cuuint64_t)M * 128};
    cuuint32_t box[3] = {64, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = tmap_encode_fn()(&ep.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dE, gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("3-D tensor map failed: %d\n", (int)r);
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int bstat = 0; bstat < 2; ++bstat) {
    auto go = [&]() {
      const unsigned long long apol = MODE == 13 ? kL2EvictLast : 0ull;
      if (bstat) launch_gemm<256, false, false, EpiProbe<MODE, WARPS>, true>(0, dA, K, dB, K, M, N, K, 1, ep, nullptr, 0, apol, false, false, g_prefetch);
      else launch_gemm<256, false, false, EpiProbe<MODE, WARPS>, false>(0, dA, K, dB, K, M, N, K, 1, ep, nullptr, 0, apol);
    };
    for (int it = 0; it < 3; ++it) go();
    CK(cudaDeviceSynchronize());
    const int iters = 10;
    cudaEventRecord(e0);
    for (int it = 0; it < iters; ++it) go();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    printf("[probe mode %d warps %2d bstat %d] %.3f ms  %.1f TFLOP/s\n", MODE, WARPS, bstat, ms, 2.0 * M * N * K / ms * 1e-9);
    {  // who waits for whom: one traced launch
      long long* tr;
      CK(cudaMalloc(&tr, 148 * 4 * sizeof(long long)));
      CK(cudaMemset(tr, 0, 148 * 4 * sizeof(long long)));
      gemm_trace_ptr() = tr;
      go();
      CK(cudaDeviceSynchronize());
      gemm_trace_ptr() = nullptr;
      long long h[148 * 4];
      CK(cudaMemcpy(h, tr, sizeof(h), cudaMemcpyDeviceToHost));
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      int n = 0;
      for (int b = 0; b < 148; ++b) {
        if (h[b * 4 + 1] == 0 && h[b * 4 + 0] == 0) continue;
        s0 += h[b * 4]; s1 += h[b * 4 + 1]; s2 += h[b * 4 + 2]; s3 += h[b * 4 + 3]; ++n;
      }
      if (n) printf("    wait kcycles per CTA: producer on free slot %.0f | MMA on operands %.0f | MMA on free accumulator %.0f | epilogue warp 0 on accumulator %.0f   (kernel ~%.0f kcycles at 1.9 GHz)\n",
                    s0 / n / 1e3, s1 / n / 1e3, s2 / n / 1e3, s3 / n / 1e3, ms * 1.9e6 / 1e3);
      cudaFree(tr);
    }
  }
}

template <int MODE>
static void probe_two_cta(const void* dA, const void* dB, void* dE, float* sink, int M, int N, int K) {
  typename EpiProbe<MODE, 8>::Params ep;
  memset(&ep, 0, sizeof(ep));
  ep.sink = sink;
  make_store_tmap_bf16(&ep.tm, dE, M, N, N);
  if (MODE == 11) make_store_tmap_bf16_slab(&ep.tm, dE, M, N);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) launch_gemm2_bstat<false, EpiProbe<MODE, 8>>(0, dA, K, dB, K, M, N, K, ep);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int it = 0; it < 10; ++it) launch_gemm2_bstat<false, EpiProbe<MODE, 8>>(0, dA, K, dB, K, M, N, K, ep);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 10;
  printf("[probe mode %d 2-CTA] %.3f ms  %.1f TFLOP/s\n", MODE, ms, 2.0 * M * N * K / ms * 1e-9);
  long long* tr;
  CK(cudaMalloc(&tr, 148 * 4 * sizeof(long long)));
  CK(cudaMemset(tr, 0, 148 * 4 * sizeof(long long)));
  gemm_trace_ptr() = tr;
  launch_gemm2_bstat<false, EpiProbe<MODE, 8>>(0, dA, K, dB, K, M, N, K, ep);
  CK(cudaDeviceSynchronize());
  gemm_trace_ptr() = nullptr;
  long long h[148 * 4];
  CK(cudaMemcpy(h, tr, sizeof(h), cudaMemcpyDeviceToHost));
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  int n = 0, nl = 0;
  for (int b = 0; b < 144; ++b) {
    s0 += h[b * 4]; s3 += h[b * 4 + 3]; ++n;
    if ((b & 1) == 0) { s1 += h[b * 4 + 1]; s2 += h[b * 4 + 2]; ++nl; }
  }
  printf("    wait kcycles per CTA: producer on free slot %.0f | leader MMA on operands %.0f | leader MMA on free accumulator %.0f | epilogue warp 0 on accumulator %.0f   (kernel ~%.0f kcycles at 1.9 GHz)\n",
         s0 / n / 1e3, s1 / nl / 1e3, s2 / nl / 1e3, s3 / n / 1e3, ms * 1.9e6 / 1e3);
  cudaFree(tr);
}

static int perf_probe() {
  const int M = 200704, N = 2048, K = 256;
  void *dA, *dB, *dE;
  float* sink;
  CK(cudaMalloc(&dA, (size_t)M * K * 2));
  CK(cudaMalloc(&dB, (size_t)N * K * 2));
  CK(cudaMalloc(&dE, (size_t)M * N * 2 + 4096));
  CK(cudaMalloc(&sink, 64));
  CK(cudaMemset(dA, 0x3c, (size_t)M * K * 2));
  CK(cudaMemset(dB, 0x3c, (size_t)N * K * 2));
  probe_two_cta<0>(dA, dB, dE, sink, M, N, K);
  probe_two_cta<2>(dA, dB, dE, sink, M, N, K);
  probe_two_cta<4>(dA, dB, dE, sink, M, N, K);
  probe_two_cta<6>(dA, dB, dE, sink, M, N, K);
  probe_two_cta<11>(dA, dB, dE, sink, M, N, K);
  probe_one<0, 8>(dA, dB, dE, sink, M, N, K);
  probe_one<2, 8>(dA, dB, dE, sink, M, N, K);
  probe_one<11, 8>(dA, dB, dE, sink, M, N, K);
  probe_one<16, 8>(dA, dB, dE, sink, M, N, K);
  probe_one<14, 8>(dA, dB, dE, (float*)dE, M, N, K);
  probe_one<15, 8>(dA, dB, dE, (float*)dE, M, N, K);
  for (int pf : {2, 4}) {
    g_prefetch = pf;
    printf("-- A tiles L2-prefetched %d steps ahead\n", pf);
    probe_one<11, 8>(dA, dB, dE, sink, M, N, K);
  }
  g_prefetch = 0;
  for (int sk : {2, 4, 8}) {
    gemm_a_skip() = sk;
    printf("-- only every %d-th A stage is actually loaded (what a %d-CTA multicast would leave per SM)\n", sk, sk);
    probe_one<11, 8>(dA, dB, dE, sink, M, N, K);
    probe_one<0, 8>(dA, dB, dE, sink, M, N, K);
  }
  gemm_a_skip() = 0;
  probe_one<12, 8>(dA, dB, dE, sink, M, N, K);
  probe_one<13, 8>(dA, dB, dE, sink, M, N, K);
  return 0;
}

// Write-path microbenchmarks: what does a pure 822 MB output stream cost without any GEMM in front of it?
__global__ void wr_plain_kernel(uint4* out, size_t n16) {
  const uint4 q = make_uint4(1, 2, 3, 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) out[i] = q;
}
// every warp stores 32-row x 64-col bf16 slabs (like the GEMM epilogue) or, ROWS=128, one 128 x 64 slab per CTA pass
template <int ROWS>
__global__ void wr_tma_kernel(const __grid_constant__ CUtensorMap tm, int M, int N) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = blockDim.x / 32;
  const int col_slabs = N / 64;
  if (ROWS == 32) {
    const long long total = (long long)(M / 32) * col_slabs;
    for (long long s = (long long)blockIdx.x * nw + warp; s < total; s += (long long)gridDim.x * nw) {
      const int r = (int)(s / col_slabs), c = (int)(s % col_slabs);
      if (lane == 0) {
        bulk_wait_read<0>();
        tma_store_2d(&tm, sm + warp * 4096, c * 64, r * 32);
        bulk_commit();
      }
      __syncwarp();
    }
  } else {
    const long long total = (long long)(M / ROWS) * col_slabs;
    for (long long s = blockIdx.x; s < total; s += gridDim.x) {
      const int r = (int)(s / col_slabs), c = (int)(s % col_slabs);
      if (threadIdx.x == 0) {
        bulk_wait_read<1>();
        tma_store_2d(&tm, sm + (s & 1) * ROWS * 128, c * 64, r * ROWS);
        bulk_commit();
      }
    }
  }
  if (lane == 0) bulk_wait<0>();
}
static int make_tmap_rows(CUtensorMap* out, void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  return make_tmap_bf16_2d(out, ptr, rows, cols, cols, box_rows);
}
static int perf_write() {
  const int M = 200704, N = 2048;
  void* dE;
  CK(cudaMalloc(&dE, (size_t)M * N * 2));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const size_t bytes = (size_t)M * N * 2;
  auto timeit = [&](const char* name, auto&& fn) {
    for (int it = 0; it < 2; ++it) fn();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int it = 0; it < 10; ++it) fn();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 10;
    printf("[write %-28s] %.3f ms  %.1f GB/s\n", name, ms, bytes / ms * 1e-6);
  };
  timeit("cudaMemsetAsync", [&]() { cudaMemsetAsync(dE, 1, bytes, 0); });
  timeit("st.global.v4 148x8x256", [&]() { wr_plain_kernel<<<148 * 8, 256>>>((uint4*)dE, bytes / 16); });
  timeit("st.global.v4 148x2x1024", [&]() { wr_plain_kernel<<<148 * 2, 1024>>>((uint4*)dE, bytes / 16); });
  CUtensorMap tm32, tm128, tm256;
  make_tmap_rows(&tm32, dE, M, N, 32);
  make_tmap_rows(&tm128, dE, M, N, 128);
  make_tmap_rows(&tm256, dE, M, N, 256);
  cudaFuncSetAttribute(wr_tma_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(wr_tma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(wr_tma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  timeit("TMA 32x64 slabs, 8 warps/CTA", [&]() { wr_tma_kernel<32><<<148, 256, 32768>>>(tm32, M, N); });
  timeit("TMA 32x64 slabs, 16 warps/CTA", [&]() { wr_tma_kernel<32><<<148, 512, 65536>>>(tm32, M, N); });
  timeit("TMA 128x64 slabs, 1/CTA", [&]() { wr_tma_kernel<128><<<148, 32, 32768>>>(tm128, M, N); });
  timeit("TMA 256x64 slabs, 1/CTA", [&]() { wr_tma_kernel<256><<<148, 32, 65536>>>(tm256, M, N); });
  timeit("TMA 128x64 slabs, 2 CTA/SM", [&]() { wr_tma_kernel<128><<<296, 32, 32768>>>(tm128, M, N); });
  {  // same byte count into a 16 MB window (L2-resident): the SM-side store rate without HBM behind it
    CUtensorMap tmw;
    make_tmap_rows(&tmw, dE, 4096, N, 32);
    timeit("TMA 32x64 slabs, 16 MB window x49", [&]() { for (int r = 0; r < 7; ++r) wr_tma_kernel<32><<<148, 256, 32768>>>(tmw, 4096, N); });
  }
  return 0;
}

// dE-shaped GEMM at cfg2 (A K-major [T,C], B MN-major [C,F], K = C = 256) with the three column-sum strategies
template <class Epi>
static void perf_de_one(const char* name, const void* dA, const void* dB, void* dE, uint32_t* mask, float* cs, int M, int N, int K) {
  typename Epi::Params ep;
  memset(&ep, 0, sizeof(ep));
  ep.mask_words = mask; ep.words = N / 32; ep.colsum_partial = cs; ep.l1c = 0.5f;
  make_store_tmap_bf16_chunk(&ep.tm_dpre, dE, M, N, N);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) launch_gemm<256, false, true, Epi, true>(0, dA, K, dB, N, M, N, K, 1, ep);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int it = 0; it < 10; ++it) launch_gemm<256, false, true, Epi, true>(0, dA, K, dB, N, M, N, K, 1, ep);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 10;
  printf("[perf dE %-22s] %.3f ms  %.1f TFLOP/s\n", name, ms, 2.0 * M * N * K / ms * 1e-9);
}
static int perf_de() {
  const int M = 200704, N = 2048, K = 256;
  void *dA, *dB, *dE;
  uint32_t* mask;
  float* cs;
  CK(cudaMalloc(&dA, (size_t)M * K * 2));
  CK(cudaMalloc(&dB, (size_t)N * K * 2));
  CK(cudaMalloc(&dE, (size_t)M * N * 2));
  CK(cudaMalloc(&mask, (size_t)M * (N / 32) * 4));
  CK(cudaMalloc(&cs, (size_t)(M / 32 + 4) * N * 4));
  CK(cudaMemset(dA, 0x3c, (size_t)M * K * 2));
  CK(cudaMemset(dB, 0x3c, (size_t)N * K * 2));
  CK(cudaMemset(mask, 0x5a, (size_t)M * (N / 32) * 4));
  perf_de_one<EpiDPre>("shuffle per tile", dA, dB, dE, mask, cs, M, N, K);
  perf_de_one<EpiDPreCta>("per-thread accumulators", dA, dB, dE, mask, cs, M, N, K);
  perf_de_one<EpiDPreNoSum>("no column sums", dA, dB, dE, mask, cs, M, N, K);
  return 0;
}

// decoder-shaped GEMM at cfg2 (M = T, N = C = 256, K = F = 2048): how much does the operand-ring depth matter?
// EpiStore has 34 KB of epilogue smem (4 stages of 48 KB); EpiStorePad pretends to need 66 KB like EpiDec (3 stages).
struct EpiStorePad : EpiStore {
  static constexpr uint32_t kSmemBytes = 2 * SlabWriter1::bytes(kWarps) + 2 * 256 * sizeof(float);
  using EpiStore::EpiStore;
};
template <class Epi>
static void perf_dec_one(const char* name, const void* dA, const void* dB, void* dD, float* bias, int M, int N, int K) {
  EpiStore::Params ep;
  memset(&ep, 0, sizeof(ep));
  ep.out = dD; ep.ld = N; ep.bias = bias; ep.alpha = 1.f; ep.relu = 0; ep.out_bf16 = 1;
  if (make_store_tmap_bf16(&ep.tm, dD, M, N, N) == 0) ep.tm_valid = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) launch_gemm<256, false, false, Epi>(0, dA, K, dB, K, M, N, K, 1, ep);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int it = 0; it < 10; ++it) launch_gemm<256, false, false, Epi>(0, dA, K, dB, K, M, N, K, 1, ep);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 10;
  printf("[perf dec %-18s stages=%d] %.3f ms  %.1f TFLOP/s  A stream %.1f GB/s\n", name,
         GemmCfg<256, Epi::kSmemBytes>::kStages, ms, 2.0 * M * N * K / ms * 1e-9, (double)M * K * 2 / ms * 1e-6);
}
static int perf_dec() {
  const int M = 200704, N = 256, K = 2048;
  void *dA, *dB, *dD;
  float* bias;
  CK(cudaMalloc(&dA, (size_t)M * K * 2));
  CK(cudaMalloc(&dB, (size_t)N * K * 2));
  CK(cudaMalloc(&dD, (size_t)M * N * 2));
  CK(cudaMalloc(&bias, N * 4));
  CK(cudaMemset(dA, 0x3c, (size_t)M * K * 2));
  CK(cudaMemset(dB, 0x3c, (size_t)N * K * 2));
  CK(cudaMemset(bias, 0, N * 4));
  perf_dec_one<EpiStore>("EpiStore", dA, dB, dD, bias, M, N, K);
  perf_dec_one<EpiStorePad>("EpiStore padded", dA, dB, dD, bias, M, N, K);
  {  // the same GEMM on SM pairs (streaming two-CTA kernel): each CTA loads half of every W_dec k-block
    EpiStore::Params ep;
    memset(&ep, 0, sizeof(ep));
    ep.out = dD; ep.ld = N; ep.bias = bias; ep.alpha = 1.f; ep.relu = 0; ep.out_bf16 = 1;
    if (make_store_tmap_bf16(&ep.tm, dD, M, N, N) == 0) ep.tm_valid = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rev = 0; rev < 2; ++rev) {
      for (int it = 0; it < 3; ++it) launch_gemm2_stream<false, false, EpiStore>(0, dA, K, dB, K, M, N, K, 1, ep, nullptr, 0, false, false, rev);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      for (int it = 0; it < 10; ++it) launch_gemm2_stream<false, false, EpiStore>(0, dA, K, dB, K, M, N, K, 1, ep, nullptr, 0, false, false, rev);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= 10;
      printf("[perf dec 2-CTA streaming EpiStore stages=%d reverse_m=%d] %.3f ms  %.1f TFLOP/s  A stream %.1f GB/s\n",
             Gemm2StreamCfg<EpiStore::kSmemBytes>::kStages, rev, ms, 2.0 * M * N * K / ms * 1e-9, (double)M * K * 2 / ms * 1e-6);
    }
  }
  {  // dW_dec shape: DIFF^T [256 x T] * E [T x 2048], split-K, fp32 partials; single-CTA against pairs
    const int M2 = 256, N2 = 2048, K2 = 200704;
    float* dP;
    CK(cudaMalloc(&dP, (size_t)16 * M2 * N2 * 4));
    EpiPartial::Params ep{dP, N2, (long long)M2 * N2};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int two = 0; two < 2; ++two) {
      int used = 0;
      auto launch = [&]() {
        if (two) launch_gemm2_stream<true, true, EpiPartial>(0, dD, M2, dA, N2, M2, N2, K2, 0, ep, &used);
        else launch_gemm<256, true, true, EpiPartial>(0, dD, M2, dA, N2, M2, N2, K2, 0, ep, &used);
      };
      for (int it = 0; it < 3; ++it) launch();
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      for (int it = 0; it < 10; ++it) launch();
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= 10;
      printf("[perf dWdec %s splits=%d] %.3f ms  %.1f TFLOP/s  B stream %.1f GB/s\n", two ? "2-CTA streaming" : "1-CTA", used, ms,
             2.0 * M2 * N2 * K2 / ms * 1e-9, (double)N2 * K2 * 2 / ms * 1e-6);
    }
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
  if (std::string(argv[1]) == "probe") return perf_probe();
  if (std::string(argv[1]) == "write") return perf_write();
  if (std::string(argv[1]) == "perf_de") return perf_de();
  if (std::string(argv[1]) == "perf_dec") return perf_dec();
  const int i = atoi(argv[1]);
  if (i < 0 || i >= (int)(sizeof(kCases) / sizeof(kCases[0]))) return 3;
  return dispatch(kCases[i]);
}
