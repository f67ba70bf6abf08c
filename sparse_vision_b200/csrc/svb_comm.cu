// Data-parallel exchange of the flat step buffer over NVLink peer memory.
//
// torch.distributed / NCCL is the plumbing (rendezvous, the 64-byte IPC handles travel through it), but the exchange
// itself is ONE kernel of ours: the flat buffer of svb_*_step_grads lives in a cudaMalloc'ed region that every rank of
// the node maps with CUDA IPC, and a two-shot all-reduce runs directly on it --
//   1. every CTA tells its twin on every peer that the local buffer is complete and waits for theirs (flags in peer
//      memory, release / acquire at system scope);
//   2. rank r reduces slice r of the SUM section (and of the MAX section) by loading it from all ranks over NVLink,
//      ALWAYS in rank order 0..world-1, and stores the result into every rank's buffer;
//   3. a second flag round makes sure all slices have landed before the optimiser half of the step reads them.
// Compared with two NCCL calls per step (~55 us at 2 GPUs, almost all launch / protocol latency for 4 MB) this is one
// launch, the MAX section rides along, and the result is bit-identical on every rank and from run to run.
//
// NVLS: when the buffer is SYMMETRIC memory with a multicast mapping (svb_comm_attach: the caller -- parallel.py through
// torch.distributed._symmetric_memory, which does the VMM allocation, the file-descriptor exchange and the multicast
// binding -- hands over every rank's pointer and the multicast pointer), step 2 of the SUM section is one
// multimem.ld_reduce (the NVSwitch adds the W replicas) and one multimem.st (the switch broadcasts the sum) per 16
// bytes instead of W loads and W stores over NVLink.
#include <cstdlib>
#include "svb_common.cuh"

using namespace svb;

namespace {

constexpr int kCommBlocks = 128, kCommThreads = 512, kMaxRanks = 8;

struct CommArgs {
  float* buf[kMaxRanks];      // every rank's flat buffer (own entry = local pointer)
  float* mc;                  // multicast mapping of the same buffer on all ranks (NVLS), or null
  uint32_t* flag[kMaxRanks];  // every rank's flag array [2 rounds][kCommBlocks][kMaxRanks]
  uint32_t* status;           // LOCAL status word: 0 = fine, else 1 + the peer that never arrived (sticky, host-readable)
  long long n_sum, n_max;
  unsigned long long timeout_ns;
  int rank, world;
  const uint32_t* epoch_ptr;  // LOCAL device word holding the number of this exchange (bumped by a one-thread kernel in
                              // front of the launch: nothing call-dependent in the launch parameters -> CUDA-graph capture)
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Round `round` of the cross-GPU rendezvous of CTA b: signal every peer's twin CTA, then wait for all of them.
// Ranks may legitimately be seconds apart (data loading, rank-0-only checkpointing), so the wait is long
// (svb_comm_set_timeout / SVB_COMM_TIMEOUT_S, default 120 s of %globaltimer).  A peer that never arrives must end
// neither as a hung GPU nor as a poisoned context: the launch gives up, records who was missing in the sticky status
// word (svb_comm_status reads it on the host; the step's results are then undefined) and every other wait of this and
// of later launches returns at once.
__device__ __forceinline__ void cta_rendezvous(const CommArgs& a, int round) {
  __syncthreads();
  if (threadIdx.x < a.world) {
    const int peer = threadIdx.x;
    const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(a.epoch_ptr);
    __threadfence_system();
    st_release_sys(a.flag[peer] + (static_cast<size_t>(round) * kCommBlocks + blockIdx.x) * kMaxRanks + a.rank, epoch);
    const uint32_t* mine = a.flag[a.rank] + (static_cast<size_t>(round) * kCommBlocks + blockIdx.x) * kMaxRanks + peer;
    const unsigned long long t0 = globaltimer_ns();
    unsigned spins = 0;
    while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) {
      if ((++spins & 1023u) == 0) {
        if (*reinterpret_cast<volatile uint32_t*>(a.status) != 0) break;
        if (globaltimer_ns() - t0 > a.timeout_ns) {
          atomicCAS(a.status, 0u, 1u + static_cast<uint32_t>(peer));
          break;
        }
      }
    }
  }
  __syncthreads();
}

template <bool MAX>
__device__ __forceinline__ float4 combine(float4 a, const float4 v) {
  if (MAX) { a.x = fmaxf(a.x, v.x); a.y = fmaxf(a.y, v.y); a.z = fmaxf(a.z, v.z); a.w = fmaxf(a.w, v.w); }
  else { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
  return a;
}
// W = compile-time bound on the world size (2, 4 or 8): 16 / W elements per thread and iteration, so that 16
// independent 16-byte loads over NVLink are in flight per thread whatever the world size (64 registers).
template <bool MAX, int W>
__device__ __forceinline__ void reduce_section(const CommArgs& a, long long begin, long long n) {
  // slice of this rank, in float4 units so that every access is 16 bytes (section starts are 16-byte aligned)
  const long long n4 = (n + 3) / 4;
  const long long per = (n4 + a.world - 1) / a.world;
  const long long lo = a.rank * per, hi = min(n4, lo + per);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  constexpr int kUnroll = 16 / W;
  for (long long i = lo + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < hi; i += kUnroll * stride) {
    float4 v[kUnroll][W];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (i + u * stride < hi) {
        const long long e = begin + 4 * (i + u * stride);
#pragma unroll
        for (int r = 0; r < W; ++r)
          if (r < a.world) v[u][r] = *reinterpret_cast<const float4*>(a.buf[r] + e);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      if (i + u * stride < hi) {
        const long long e = begin + 4 * (i + u * stride);
        float4 acc = v[u][0];   // always in rank order 0..world-1: bit-identical on every rank
#pragma unroll
        for (int r = 1; r < W; ++r)
          if (r < a.world) acc = combine<MAX>(acc, v[u][r]);
#pragma unroll
        for (int r = 0; r < W; ++r)
          if (r < a.world) *reinterpret_cast<float4*>(a.buf[r] + e) = acc;
      }
    }
  }
}

// SUM section through the switch: rank r owns slice r; multimem.ld_reduce returns the sum of the W replicas of 16 bytes,
// multimem.st writes it to all of them.  The same value lands on every rank, so replicas stay bit-identical.
__device__ __forceinline__ void reduce_sum_nvls(const CommArgs& a, long long n) {
  const long long n4 = (n + 3) / 4;
  const long long per = (n4 + a.world - 1) / a.world;
  const long long lo = a.rank * per, hi = min(n4, lo + per);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  constexpr int kUnroll = 8;
  for (long long i = lo + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < hi; i += kUnroll * stride) {
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (i + u * stride < hi)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                     : "l"(a.mc + 4 * (i + u * stride))
                     : "memory");
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (i + u * stride < hi)
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.mc + 4 * (i + u * stride)),
                     "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w)
                     : "memory");
  }
}

template <int W>
__device__ __forceinline__ void allreduce_body(const CommArgs& a) {
  cta_rendezvous(a, 0);                                 // every rank's buffer is complete
  if (a.mc) reduce_sum_nvls(a, a.n_sum);
  else reduce_section<false, W>(a, 0, a.n_sum);
  reduce_section<true, W>(a, (a.n_sum + 3) / 4 * 4, a.n_max);
  cta_rendezvous(a, 1);                                 // every slice has landed everywhere
}

__global__ void __launch_bounds__(kCommThreads) allreduce_flat_kernel(const CommArgs a) {
  if (a.world <= 2) allreduce_body<2>(a);
  else if (a.world <= 4) allreduce_body<4>(a);
  else allreduce_body<8>(a);
}

}  // namespace

struct svb_comm {
  double timeout_s = 120.0;
  int rank = 0, world = 1;
  float* base = nullptr;         // local region: [capacity floats | flags]
  int64_t capacity = 0;          // floats
  void* peer_base[kMaxRanks] = {nullptr};
  float* mc = nullptr;           // multicast mapping (NVLS) or null
  bool owned = true;             // false: the region belongs to the caller (svb_comm_attach)
  bool connected = false;
};

// flags of the two rendezvous rounds, then 64 bytes whose first word is the status
static size_t comm_flag_bytes() { return sizeof(uint32_t) * 2 * kCommBlocks * kMaxRanks + 64; }
static uint32_t* comm_status_word(svb_comm* c) {
  return reinterpret_cast<uint32_t*>(c->base + c->capacity) + 2 * kCommBlocks * kMaxRanks;
}
// the exchange counter lives beside the status word (zeroed with the region; every rank counts the same exchanges)
static uint32_t* comm_epoch_word(svb_comm* c) { return comm_status_word(c) + 1; }
namespace {
__global__ void comm_epoch_bump_kernel(uint32_t* e) { *e += 1; }
}

extern "C" int svb_comm_alloc(svb_handle* h, int64_t n_floats, void* ipc_handle_out) {
  if (!h || n_floats <= 0 || !ipc_handle_out) return fail(SVB_ERR_BAD_ARG, "svb_comm_alloc: bad argument");
  SVB_ON_DEVICE(h);
  if (h->comm_ctx) return fail(SVB_ERR_BAD_ARG, "svb_comm_alloc: a communication buffer already exists");
  svb_comm* c = new svb_comm();
  if (const char* env = getenv("SVB_COMM_TIMEOUT_S")) {
    const double v = atof(env);
    if (v > 0) c->timeout_s = v;
  }
  c->capacity = (n_floats + 63) / 64 * 64;
  const size_t bytes = static_cast<size_t>(c->capacity) * 4 + comm_flag_bytes();
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    delete c;
    return fail(SVB_ERR_NOMEM, "svb_comm_alloc: cudaMalloc of %zu bytes failed", bytes);
  }
  cudaMemset(p, 0, bytes);
  cudaDeviceSynchronize();
  c->base = static_cast<float*>(p);
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaFree(p);
    delete c;
    return fail(SVB_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(ipc_handle_out, &hd, 64);
  h->comm_ctx = c;
  return 0;
}

extern "C" int svb_comm_connect(svb_handle* h, int32_t rank, int32_t world, const void* ipc_handles) {
  if (!h || !h->comm_ctx || !ipc_handles) return fail(SVB_ERR_BAD_ARG, "svb_comm_connect: call svb_comm_alloc first");
  SVB_ON_DEVICE(h);
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world)
    return fail(SVB_ERR_UNSUPPORTED, "svb_comm_connect: world size %d (1..%d ranks of one node)", world, kMaxRanks);
  svb_comm* c = h->comm_ctx;
  c->rank = rank;
  c->world = world;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { c->peer_base[r] = c->base; continue; }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, static_cast<const char*>(ipc_handles) + 64 * r, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(SVB_ERR_CUDA, "cudaIpcOpenMemHandle for rank %d failed: %s", r, cudaGetErrorString(e));
    }
    c->peer_base[r] = p;
  }
  c->connected = true;
  return 0;
}

extern "C" int svb_comm_region_bytes(int64_t n_floats, int64_t* bytes, int64_t* capacity_floats) {
  if (n_floats <= 0 || !bytes) return fail(SVB_ERR_BAD_ARG, "svb_comm_region_bytes: bad argument");
  const int64_t cap = (n_floats + 63) / 64 * 64;
  *bytes = cap * 4 + static_cast<int64_t>(comm_flag_bytes());
  if (capacity_floats) *capacity_floats = cap;
  return 0;
}

extern "C" int svb_comm_attach(svb_handle* h, int32_t rank, int32_t world, int64_t n_floats, const void* const* region_ptrs_host,
                               void* multicast_ptr) {
  if (!h || !region_ptrs_host || n_floats <= 0) return fail(SVB_ERR_BAD_ARG, "svb_comm_attach: bad argument");
  SVB_ON_DEVICE(h);
  if (h->comm_ctx) return fail(SVB_ERR_BAD_ARG, "svb_comm_attach: a communication buffer already exists");
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world)
    return fail(SVB_ERR_UNSUPPORTED, "svb_comm_attach: world size %d (1..%d ranks of one node)", world, kMaxRanks);
  for (int r = 0; r < world; ++r)
    if (!region_ptrs_host[r] || (reinterpret_cast<uintptr_t>(region_ptrs_host[r]) & 15))
      return fail(SVB_ERR_BAD_ARG, "svb_comm_attach: region pointer of rank %d is null or not 16-byte aligned", r);
  if (reinterpret_cast<uintptr_t>(multicast_ptr) & 15) return fail(SVB_ERR_BAD_ARG, "svb_comm_attach: misaligned multicast pointer");
  svb_comm* c = new svb_comm();
  if (const char* env = getenv("SVB_COMM_TIMEOUT_S")) {
    const double v = atof(env);
    if (v > 0) c->timeout_s = v;
  }
  c->capacity = (n_floats + 63) / 64 * 64;
  c->rank = rank; c->world = world;
  for (int r = 0; r < world; ++r) c->peer_base[r] = const_cast<void*>(region_ptrs_host[r]);
  c->base = static_cast<float*>(c->peer_base[rank]);
  c->mc = static_cast<float*>(multicast_ptr);
  c->owned = false;
  c->connected = true;
  h->comm_ctx = c;
  return 0;
}

extern "C" int svb_comm_capacity(svb_handle* h, int64_t* n_floats) {
  if (!h || !n_floats) return fail(SVB_ERR_BAD_ARG, "null argument");
  *n_floats = h->comm_ctx ? h->comm_ctx->capacity : 0;
  return 0;
}

extern "C" int svb_comm_allreduce(svb_handle* h, void* stream) {
  if (!h || !h->comm_ctx || !h->comm_ctx->connected) return fail(SVB_ERR_BAD_ARG, "svb_comm_allreduce: not connected");
  SVB_ON_DEVICE(h);
  svb_comm* c = h->comm_ctx;
  if (!h->gradbuf || h->gradbuf != c->base)
    return fail(SVB_ERR_BAD_ARG, "svb_comm_allreduce: the last svb_*_step_grads did not use the communication buffer");
  CommArgs a{};
  for (int r = 0; r < c->world; ++r) {
    a.buf[r] = static_cast<float*>(c->peer_base[r]);
    a.flag[r] = reinterpret_cast<uint32_t*>(static_cast<float*>(c->peer_base[r]) + c->capacity);
  }
  a.mc = c->mc;
  a.n_sum = h->sum_elems; a.n_max = h->max_elems; a.rank = c->rank; a.world = c->world;
  a.status = comm_status_word(c);
  a.timeout_ns = static_cast<unsigned long long>(c->timeout_s * 1e9);
  a.epoch_ptr = comm_epoch_word(c);
  (comm_epoch_bump_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(comm_epoch_word(c)), svb::count_launch());
  // The twin-CTA rendezvous needs all kCommBlocks CTAs of every rank on the machine at the same time: a cooperative
  // launch guarantees exactly that (the grid starts only when all of it fits; it is refused if it never can).
  static int coop_ok = -1;
  if (coop_ok < 0) {
    int coop = 0, per_sm = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, allreduce_flat_kernel, kCommThreads, 0);
    coop_ok = (coop && per_sm * h->sms >= kCommBlocks) ? 1 : 0;
  }
  if (!coop_ok)
    return fail(SVB_ERR_UNSUPPORTED, "svb_comm_allreduce: %d co-resident CTAs are not available on this device", kCommBlocks);
  void* args[] = {&a};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(allreduce_flat_kernel), dim3(kCommBlocks),
                                              dim3(kCommThreads), args, 0, static_cast<cudaStream_t>(stream));
  svb::count_launch();
  if (e != cudaSuccess) return fail(SVB_ERR_CUDA, "cooperative launch of the all-reduce failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int svb_comm_set_timeout(svb_handle* h, double seconds) {
  if (!h || !h->comm_ctx || !(seconds > 0)) return fail(SVB_ERR_BAD_ARG, "svb_comm_set_timeout: no buffer / bad value");
  h->comm_ctx->timeout_s = seconds;
  return 0;
}

// 0 = every all-reduce so far completed; 1 + r = rank r never arrived at some rendezvous within the timeout (sticky: the
// parameters of that step and of every later one are undefined).  Synchronises with the device.
extern "C" int svb_comm_status(svb_handle* h, int32_t* status) {
  if (!h || !status) return fail(SVB_ERR_BAD_ARG, "null argument");
  *status = 0;
  if (!h->comm_ctx) return 0;
  SVB_ON_DEVICE(h);
  uint32_t v = 0;
  SVB_CUDA(cudaMemcpy(&v, comm_status_word(h->comm_ctx), 4, cudaMemcpyDeviceToHost));
  *status = static_cast<int32_t>(v);
  return 0;
}

extern "C" int svb_comm_destroy(svb_handle* h) {
  if (!h || !h->comm_ctx) return 0;
  svb_comm* c = h->comm_ctx;
  cudaDeviceSynchronize();
  if (c->owned) {
    for (int r = 0; r < c->world; ++r)
      if (r != c->rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
    if (c->base) cudaFree(c->base);
  }
  delete c;
  h->comm_ctx = nullptr;
  return 0;
}

// The flat buffer a step should use: the communication buffer when it exists and is large enough, else `arena_flat`.
float* svb::comm_flat_or(svb_handle* h, float* arena_flat, size_t need_floats) {
  svb_comm* c = h->comm_ctx;
  // the MAX section starts at the next multiple of 4 after the SUM section in the exchange kernel; steps lay the two
  // sections out back to back, so only buffers whose SUM section is a multiple of 4 long qualify (always true:
  // C and F are multiples of 8)
  if (c && static_cast<size_t>(c->capacity) >= need_floats) return c->base;
  return arena_flat;
}
