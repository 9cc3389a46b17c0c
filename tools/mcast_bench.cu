// mcast_bench.cu — does TMA multicast across the CTAs that share an operand raise the delivered L2->SM bandwidth?
//
// The 384-wide GEMMs read every A tile three times (once per 128-column block; CTAs 3j, 3j+1, 3j+2 walk the same row tiles) and
// are bound by the L2->SM path.  This tool streams a 51 MB buffer through shared-memory rings with cp.async.bulk,
// no math, in three ways and reports the bytes DELIVERED per clock per SM:
//   single    every CTA streams its own chunks (no redundancy)                              -> per-SM ingest ceiling
//   unicast3  CTAs 3j..3j+2 stream the SAME chunks, each with its own copies                -> what the GEMMs do today
//   mcast3    clusters of 3: every CTA fetches a third of each chunk and multicasts it       -> each byte leaves L2 once
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/mcast_bench tools/mcast_bench.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

constexpr int kChunk = 24576;   // bytes per ring slot (three 8 KB slices)
constexpr int kSlice = kChunk / 3;
constexpr long long kSpinLimit = 4LL * 1000 * 1000 * 1000;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && clock64() - t0 > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load_mcast(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar), "h"(mask)
               : "memory");
}

// MODE 0 single, 1 unicast3, 2 mcast3 (launched with cluster dimension 3)
// DELAY: cycles the consumer holds a slot before handing it back (0 = pure streaming; 384 = the time the tensor pipe needs for the
// MMAs of 24 KB of A at 128x128x16 / 64 cycles, i.e. a consumption limit of 64 B/clk/SM)
template <int MODE, int kStages>
__global__ void __launch_bounds__(64, 1) stream_kernel(const uint8_t* __restrict__ src, long long total_chunks, long long* out, int delay) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kChunk);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kStages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = MODE == 2 ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, MODE == 2 ? 3 : 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (MODE == 2) cluster_sync_all();

  // chunk list of this CTA: MODE 0: b, b + grid, ...; MODE 1/2: the group's list g, g + G, ... (all three CTAs of a group the same)
  const long long first = MODE == 0 ? blockIdx.x : blockIdx.x / 3;
  const long long step = MODE == 0 ? gridDim.x : gridDim.x / 3;
  const long long n = first < total_chunks ? (total_chunks - first + step - 1) / step : 0;
  const long long t0 = clock64();
  if (warp == 0 && lane == 0) {  // producer
    for (long long i = 0; i < n; ++i) {
      const int s = (int)(i % kStages);
      const uint32_t ph = (uint32_t)((i / kStages) & 1);
      mbar_wait(empty0 + 8 * s, ph ^ 1u);  // first pass over the ring: passes immediately
      const uint8_t* chunk = src + (first + i * step) * kChunk;
      mbar_arrive_expect_tx(full0 + 8 * s, kChunk);
      if (MODE == 2) {
        bulk_load_mcast(smem_u32(ring + (size_t)s * kChunk + crank * kSlice), chunk + crank * kSlice, kSlice, full0 + 8 * s, (uint16_t)0x7);
      } else {
        bulk_load(smem_u32(ring + (size_t)s * kChunk), chunk, kChunk, full0 + 8 * s);
      }
    }
  } else if (warp == 1 && lane == 0) {  // consumer: no math, hand the slot straight back
    for (long long i = 0; i < n; ++i) {
      const int s = (int)(i % kStages);
      const uint32_t ph = (uint32_t)((i / kStages) & 1);
      mbar_wait(full0 + 8 * s, ph);
      if (delay > 0) {
        const long long c0 = clock64();
        while (clock64() - c0 < delay) {}
      }
      if (MODE == 2) {
        for (uint32_t r = 0; r < 3; ++r) {
          if (r == crank) mbar_arrive(empty0 + 8 * s);
          else mbar_arrive_remote(mapa_rank(empty0 + 8 * s, r));
        }
      } else {
        mbar_arrive(empty0 + 8 * s);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (MODE == 2) cluster_sync_all();  // nobody exits while a peer may still multicast into it or arrive on its barriers
  if (threadIdx.x == 0) {
    out[2 * blockIdx.x] = t1 - t0;
    out[2 * blockIdx.x + 1] = n * kChunk;
  }
}

template <int MODE, int kStages>
static void run(const char* name, const uint8_t* src, long long total_chunks, long long* d_out, int grid, int delay) {
  const size_t smem = (size_t)kStages * kChunk + 2 * kStages * 8 + 64;
  auto kern = stream_kernel<MODE, kStages>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = MODE == 2 ? 3 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best_ms = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, kern, src, total_chunks, d_out, delay);
    cudaEventRecord(e1);
    if (err != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
      printf("%-9s launch failed: %s / %s\n", name, cudaGetErrorString(err), cudaGetErrorString(cudaGetLastError()));
      return;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best_ms) best_ms = ms;
  }
  long long* h = (long long*)malloc(sizeof(long long) * 2 * grid);
  cudaMemcpy(h, d_out, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
  double cyc = 0, maxc = 0, bytes = 0;
  for (int b = 0; b < grid; ++b) {
    cyc += (double)h[2 * b];
    if ((double)h[2 * b] > maxc) maxc = (double)h[2 * b];
    bytes += (double)h[2 * b + 1];
  }
  printf("%-9s stages %d (%3d KB in flight) hold %4d cyc: %8.2f us  delivered %7.1f MB  %5.1f B/clk/SM (mean CTA), %5.1f (slowest)  = %5.2f TB/s\n", name,
         kStages, kStages * kChunk / 1024, delay, best_ms * 1e3, bytes / 1e6, bytes / cyc, bytes / grid / maxc, bytes / (best_ms * 1e-3) / 1e12);
  free(h);
}

template <int kStages>
static void sweep(const uint8_t* src, long long total_chunks, long long* d_out, int delay, bool with_mcast) {
  run<0, kStages>("single", src, total_chunks, d_out, 147, delay);
  run<1, kStages>("unicast3", src, total_chunks, d_out, 147, delay);
  if (with_mcast) run<2, kStages>("mcast3", src, total_chunks, d_out, 147, delay);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d_out = nullptr;
  cudaMalloc(&d_out, sizeof(long long) * 2 * 160);
  // 51 MB: one bf16 activation tensor of the benchmark, L2-resident after the first pass; 1 GB: every chunk comes from DRAM
  for (long long mb : {51LL, 1024LL}) {
    const long long bytes = mb == 51 ? 66560LL * 384 * 2 : mb << 20;
    const long long total_chunks = bytes / kChunk;
    uint8_t* src = nullptr;
    if (cudaMalloc(&src, total_chunks * kChunk) != cudaSuccess) return 1;
    cudaMemset(src, 1, total_chunks * kChunk);
    printf("== %d SMs, buffer %.1f MB = %lld chunks of %d B (%s)\n", sms, bytes / 1e6, total_chunks, kChunk, mb == 51 ? "fits the 126 MB L2" : "DRAM-resident");
    for (int delay : {0, 384, 768}) {
      sweep<2>(src, total_chunks, d_out, delay, false);
      sweep<4>(src, total_chunks, d_out, delay, delay == 0);
      sweep<8>(src, total_chunks, d_out, delay, false);
    }
    cudaFree(src);
  }
  return 0;
}
