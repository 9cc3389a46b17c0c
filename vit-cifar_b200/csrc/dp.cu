// dp.cu — data-parallel gradient exchange fused with the optimiser, over NVLink peer memory.
//
// What the reference gets from Lightning's DDP + torch.optim.Adam (main.py:220-231, network.py:71-77):
// all-reduce(mean) of every gradient, then the same Adam update on every replica.  Here that is ONE kernel per step and GPU:
//
//   barrier A   every rank has finished its backward (its gradient buffer is complete)
//   reduce      rank r owns elements [lo_r, hi_r) of the flat buffers: g = sum over ranks (fixed order 0..W-1) of the PEERS'
//               gradient buffers, read straight over NVLink                                   (reduce-scatter)
//   Adam        update of the owned slice of p / m / v — each rank keeps the moments of its slice only up to date (ZeRO-1 style)
//   broadcast   the new fp32 parameters and their bf16 shadow are stored into EVERY rank's parameter buffers   (all-gather)
//   barrier B   every rank's stores have landed: the next forward may start
//
// The barriers are flag exchanges in peer memory (st.release.sys / ld.acquire.sys); waits are bounded and trap instead of hanging.
// All replicas receive bit-identical parameters (each element is computed once, by its owner).
//
// Peer buffers are mapped with CUDA IPC (vitb_ipc_export / vitb_ipc_open below): torch.distributed only carries the 64-byte handles.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "common.cuh"

namespace vitb {

constexpr int kDpMaxWorld = 8;

struct DpArgs {
  const float* g[kDpMaxWorld];   // every rank's flat gradient buffer (this rank's own at [rank])
  float* p[kDpMaxWorld];         // every rank's fp32 master parameters
  bf16* c[kDpMaxWorld];          // every rank's bf16 shadow (all null in fp32 check mode)
  uint32_t* flags[kDpMaxWorld];  // every rank's flag array: flags[q][r] is written by rank r, read by rank q
  float* m;                      // this rank's Adam moments (only its slice is maintained)
  float* v;
  uint32_t* sync;                // this rank's {epoch, go, done-counter, pad}
  const float* hyper_dev;        // 16 floats (vitb_adam_multi layout) in device memory, or null
  AdamHyper hyper;               // used when hyper_dev is null
  int64_t lo4, hi4;              // owned slice in float4 units
  int rank, world;
  int optimizer;                 // 0 = Adam, 1 = SGD with momentum (m is the momentum buffer, v unused)
  long long spin_limit;          // bound of every flag wait in SM clocks (vitb_dp_set_timeout / VITB_DP_TIMEOUT_S)
  long long* phase_clk;          // optional [5] per-launch phase durations of block 0 / the last block (tools/dp_phases.py); null in production
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// system-scope load: peer memory must not be served from this SM's L1
__device__ __forceinline__ float4 ld4_sys(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// Flag waits are bounded so that a dead peer becomes a launch error instead of a hung GPU.  The bound has to cover ordinary rank
// skew (rank-0-only validation or checkpoint writing, a slow data loader, first-step graph capture): default 600 s of SM clocks at
// ~2 GHz, like a collective library's watchdog; VITB_DP_TIMEOUT_S or vitb_dp_set_timeout() change it.  All ranks must enter
// step() within that time of each other.
static long long g_dp_spin_limit = -1;
static long long dp_spin_limit() {
  if (g_dp_spin_limit < 0) {
    const char* e = getenv("VITB_DP_TIMEOUT_S");
    double s = e ? atof(e) : 600.0;
    if (!(s > 0.0)) s = 600.0;
    g_dp_spin_limit = (long long)(s * 2.0e9);
  }
  return g_dp_spin_limit;
}

// thread t < world: tell rank t that this rank reached `value`, then wait until rank t has told us the same
__device__ __forceinline__ void dp_exchange(const DpArgs& a, int t, uint32_t value) {
  __threadfence_system();
  st_release_sys(a.flags[t] + a.rank, value);
  const uint32_t* mine = a.flags[a.rank] + t;
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(mine) - value) < 0) {
    if (clock64() - t0 > a.spin_limit) {
      printf("vitb dp: rank %d timed out waiting for rank %d (flag %u, want %u)\n", a.rank, t, ld_acquire_sys(mine), value);
      __trap();
    }
    __nanosleep(64);
  }
}

__global__ void __launch_bounds__(256) dp_reduce_adam_kernel(const DpArgs a) {
  __shared__ uint32_t s_epoch;
  __shared__ bool s_last;
  // Under programmatic dependent launch this grid may start before the backward kernels have finished: nothing below may read
  // the gradients, touch sync[] or tell the peers "my backward is complete" before the preceding grid's memory is visible.
  pdl_wait();
  const long long c0 = clock64();
  if (threadIdx.x == 0) s_epoch = a.sync[0] + 1;  // sync[0] is only advanced by the last block of the previous launch
  __syncthreads();
  const uint32_t epoch = s_epoch;
  // ---- barrier A: block 0 trades flags with the peers, then releases the other blocks of this grid
  if (blockIdx.x == 0) {
    if ((int)threadIdx.x < a.world) dp_exchange(a, threadIdx.x, 2 * epoch - 1);
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(a.sync + 1, epoch);
  }
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_gpu(a.sync + 1) - epoch) < 0) {
      if (clock64() - t0 > 2 * a.spin_limit) __trap();
      __nanosleep(32);
    }
  }
  __syncthreads();
  const long long c1 = clock64();

  // ---- reduce-scatter + Adam + all-gather of the owned slice
  const AdamHyper h = a.hyper_dev != nullptr ? adam_hyper_from(a.hyper_dev) : a.hyper;
  const int world = a.world;
  float* pm = a.p[a.rank];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = a.lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.hi4; i += stride) {
    float4 gq[kDpMaxWorld];
#pragma unroll
    for (int q = 0; q < kDpMaxWorld; ++q)
      if (q < world) gq[q] = ld4_sys(a.g[q] + i * 4);  // all peers' loads in flight before the first add
    float4 pp = ld4(pm + i * 4), mm = ld4(a.m + i * 4), vv = a.optimizer == 0 ? ld4(a.v + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 gg = gq[0];
#pragma unroll
    for (int q = 1; q < kDpMaxWorld; ++q)
      if (q < world) { gg.x = __fadd_rn(gg.x, gq[q].x); gg.y = __fadd_rn(gg.y, gq[q].y); gg.z = __fadd_rn(gg.z, gq[q].z); gg.w = __fadd_rn(gg.w, gq[q].w); }
    if (a.optimizer == 0) {
      adam_one(pp.x, gg.x, mm.x, vv.x, h);
      adam_one(pp.y, gg.y, mm.y, vv.y, h);
      adam_one(pp.z, gg.z, mm.z, vv.z, h);
      adam_one(pp.w, gg.w, mm.w, vv.w, h);
      st4(a.v + i * 4, vv);
    } else {
      sgd_one(pp.x, gg.x, mm.x, h);
      sgd_one(pp.y, gg.y, mm.y, h);
      sgd_one(pp.z, gg.z, mm.z, h);
      sgd_one(pp.w, gg.w, mm.w, h);
    }
    st4(a.m + i * 4, mm);
#pragma unroll
    for (int q = 0; q < kDpMaxWorld; ++q)
      if (q < world) {
        st4(a.p[q] + i * 4, pp);
        if (a.c[q] != nullptr) st4(a.c[q] + i * 4, pp);
      }
  }

  // ---- barrier B: the last block to finish (all stores of the grid are then ordered before its signal) trades flags again
  const long long c2 = clock64();
  __threadfence_system();
  __syncthreads();
  if (a.phase_clk != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {  // block 0: wait for the peers' backward, then its share of the slice
    a.phase_clk[0] = c1 - c0;
    a.phase_clk[1] = c2 - c1;
  }
  if (threadIdx.x == 0) {
    const uint32_t done = atomicAdd(a.sync + 2, 1u);
    s_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    const long long c3 = clock64();
    if ((int)threadIdx.x < a.world) dp_exchange(a, threadIdx.x, 2 * epoch);
    __syncthreads();
    if (threadIdx.x == 0) {
      a.sync[2] = 0;
      a.sync[0] = epoch;
      __threadfence();
      if (a.phase_clk != nullptr) {  // last block: whole grid (entry of this block to all stores issued), fence, barrier B
        a.phase_clk[2] = c2 - c0;
        a.phase_clk[3] = c3 - c2;
        a.phase_clk[4] = clock64() - c3;
      }
    }
  }
}

// ---- CUDA IPC plumbing -------------------------------------------------------------------------------------------------
typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

static GetRangeFn get_range_fn() {
  static GetRangeFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
  fn = (GetRangeFn)p;
  return fn;
}

static long long* g_dp_phase_clk = nullptr;  // tools only (vitb_debug_dp_phases)

static std::mutex g_ipc_mu;
static std::map<std::string, void*> g_ipc_open;  // handle bytes -> mapped base (a handle may be opened once per process)

}  // namespace vitb

using namespace vitb;

extern "C" {

/* tools only (not in vitb200.h): device buffer of 5 int64 that every dp_reduce_adam launch fills with its phase durations
   {barrier A, block 0's slice, last block's entry-to-stores, fence, barrier B} in SM clocks; NULL switches it off */
int vitb_debug_dp_phases(long long* dev_buf) {
  g_dp_phase_clk = dev_buf;
  return 0;
}

int vitb_dp_set_timeout(double seconds) {
  VITB_REQUIRE(seconds > 0.0, "dp_set_timeout: seconds must be positive");
  g_dp_spin_limit = (long long)(seconds * 2.0e9);
  return 0;
}

int vitb_ipc_export(const void* dev_ptr, void* handle64, int64_t* offset) {
  VITB_REQUIRE(dev_ptr && handle64 && offset, "ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C ABI promises 64-byte handles");
  GetRangeFn fn = get_range_fn();
  VITB_REQUIRE(fn != nullptr, "ipc_export: cuMemGetAddressRange not available from the driver");
  CUdeviceptr base = 0;
  size_t size = 0;
  const CUresult r = fn(&base, &size, (CUdeviceptr)dev_ptr);
  VITB_REQUIRE(r == CUDA_SUCCESS, "ipc_export: cuMemGetAddressRange failed with CUresult %d", (int)r);
  cudaIpcMemHandle_t h;
  VITB_CUDA_OK(cudaIpcGetMemHandle(&h, (void*)base));
  std::memcpy(handle64, &h, 64);
  *offset = (int64_t)((CUdeviceptr)dev_ptr - base);
  return 0;
}

int vitb_ipc_open(const void* handle64, int64_t offset, void** dev_ptr) {
  VITB_REQUIRE(handle64 && dev_ptr && offset >= 0, "ipc_open: bad argument");
  std::lock_guard<std::mutex> lk(g_ipc_mu);
  const std::string key((const char*)handle64, 64);
  auto it = g_ipc_open.find(key);
  void* base = nullptr;
  if (it != g_ipc_open.end()) {
    base = it->second;
  } else {
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    VITB_CUDA_OK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    g_ipc_open[key] = base;
  }
  *dev_ptr = (char*)base + offset;
  return 0;
}

int vitb_dp_reduce_adam(const void* const* g_peers, void* const* p_peers, void* const* shadow_peers, void* const* flag_peers, float* m, float* v,
                        uint32_t* sync, int64_t n, int rank, int world, int optimizer, const float* hyper_host, const float* hyper_dev, void* stream) {
  VITB_REQUIRE(g_peers && p_peers && flag_peers && m && sync && (v || optimizer == 1), "dp_reduce_adam: null pointer");
  VITB_REQUIRE(optimizer == 0 || optimizer == 1, "dp_reduce_adam: optimizer %d (0 = Adam, 1 = SGD)", optimizer);
  VITB_REQUIRE(world >= 1 && world <= kDpMaxWorld && rank >= 0 && rank < world, "dp_reduce_adam: rank %d / world %d (at most %d ranks)", rank, world,
               kDpMaxWorld);
  VITB_REQUIRE(n > 0 && n % 4 == 0, "dp_reduce_adam: the element count (%lld) must be a positive multiple of 4", (long long)n);
  VITB_REQUIRE(hyper_host || hyper_dev, "dp_reduce_adam: need hyper_host or hyper_dev");
  DpArgs a = {};
  for (int q = 0; q < world; ++q) {
    VITB_REQUIRE(g_peers[q] && p_peers[q] && flag_peers[q], "dp_reduce_adam: null peer pointer for rank %d", q);
    VITB_REQUIRE(((uintptr_t)g_peers[q] | (uintptr_t)p_peers[q]) % 16 == 0 && (shadow_peers == nullptr || (uintptr_t)shadow_peers[q] % 8 == 0),
                 "dp_reduce_adam: peer buffers must be 16-byte aligned");
    a.g[q] = (const float*)g_peers[q];
    a.p[q] = (float*)p_peers[q];
    a.c[q] = shadow_peers != nullptr ? (bf16*)shadow_peers[q] : nullptr;
    a.flags[q] = (uint32_t*)flag_peers[q];
  }
  a.m = m; a.v = v; a.sync = sync; a.hyper_dev = hyper_dev;
  if (hyper_host) a.hyper = adam_hyper_from(hyper_host);
  const int64_t n4 = n / 4, per = (n4 + world - 1) / world;
  a.lo4 = per * rank < n4 ? per * rank : n4;
  a.hi4 = per * (rank + 1) < n4 ? per * (rank + 1) : n4;
  a.rank = rank; a.world = world; a.optimizer = optimizer;
  a.spin_limit = dp_spin_limit();
  a.phase_clk = g_dp_phase_clk;
  int blocks = (int)ceil_div64(a.hi4 - a.lo4 + 1, 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  if (blocks < 1) blocks = 1;
  VITB_LAUNCH((dp_reduce_adam_kernel), blocks, 256, 0, (cudaStream_t)stream, a);
  VITB_LAUNCH_OK();
  return 0;
}

}  // extern "C"
