// tools/mma_bench.cu — microbenchmark (not part of the library): how fast can ONE thread feed tcgen05.mma, and what is the
// SS-mode execution floor for 128 x N x 16 bf16 MMAs whose operands sit in 128B-swizzled shared memory?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/mma_bench tools/mma_bench.cu
//   tools/bin/mma_bench
// Variants: 0 = descriptors rebuilt per MMA inside an `if (lane == 0)` branch (gemm_tc.cu v2 style)
//           1 = whole warp runs the loop, descriptor low words are one add away, issue under elect.sync
// Optional background traffic: one thread streams cp.async.bulk copies (L2 -> smem) like a TMA producer would.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
               "r"(idesc), "r"(accumulate)
               : "memory");
}
// A operand in tensor memory (the "TS" form): D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
  return pred != 0;
}

constexpr int STAGES = 4;
constexpr int KB_PER_TILE = 6;

template <int N, int VARIANT, bool MN_MAJOR>
__global__ void __launch_bounds__(384, 1) mma_bench_kernel(long long* out, int tiles, const uint8_t* gsrc, int traffic_bytes, int ld_mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = 128 * 64 * 2, B_BYTES = N * 64 * 2, STAGE = A_BYTES + B_BYTES;
  const uint32_t bars = base + STAGES * STAGE;           // STAGES dummy "empty" barriers + 1 final + 2 traffic
  const uint32_t traffic_dst = bars + 128;                // 2 x 8 KB landing zone
  const uint32_t tmem_slot = bars + 120;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // zero the operands (values do not matter, NaN-free keeps the data path honest)
  for (uint32_t i = threadIdx.x; i < STAGES * STAGE / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i * 4), "r"(0x3c003c00u));
  if (threadIdx.x == 0) {
    asm volatile("st.volatile.shared.b32 [%0], %1;" ::"r"(bars + 112), "r"(0));
    for (int s = 0; s < STAGES + 3; ++s) mbar_init(bars + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((MN_MAJOR ? 1u : 0u) << 15) | ((MN_MAJOR ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t final_bar = bars + STAGES * 8;

  if (warp == 1) {
    long long t0 = clock64();
    if (VARIANT == 0) {
      if (lane == 0) {
        int stage = 0;
        for (int t = 0; t < tiles; ++t) {
          const uint32_t d = tmem_base + (uint32_t)((t & 1) * N);
          for (int kb = 0; kb < KB_PER_TILE; ++kb) {
            const uint32_t a = base + stage * STAGE, b = a + A_BYTES;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = MN_MAJOR ? make_desc(a + kk * 2048, 64 * 128, 1024) : make_desc(a + kk * 32, 16, 1024);
              const uint64_t bd = MN_MAJOR ? make_desc(b + kk * 2048, 64 * 128, 1024) : make_desc(b + kk * 32, 16, 1024);
              tc_mma(d, ad, bd, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            }
            tc_commit(bars + stage * 8);
            if (++stage == STAGES) stage = 0;
          }
        }
        tc_commit(final_bar);
      }
    } else {
      // descriptor = constant high word | low word; low word = (addr >> 4) | (lbo >> 4) << 16
      const uint32_t hi = (uint32_t)((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      const uint32_t lbo_field = (MN_MAJOR ? ((64u * 128u) >> 4) : 1u) << 16;
      const uint32_t kstep = MN_MAJOR ? (2048u >> 4) : (32u >> 4);
      const bool leader = elect_one();
      int stage = 0;
      for (int t = 0; t < tiles; ++t) {
        const uint32_t d = tmem_base + (uint32_t)((t & 1) * N);
        for (int kb = 0; kb < KB_PER_TILE; ++kb) {
          const uint32_t a_lo = (((base + stage * STAGE) & 0x3FFFFu) >> 4) | lbo_field;
          const uint32_t b_lo = (((base + stage * STAGE + A_BYTES) & 0x3FFFFu) >> 4) | lbo_field;
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = ((uint64_t)hi << 32) | (uint64_t)(a_lo + kk * kstep);
              const uint64_t bd = ((uint64_t)hi << 32) | (uint64_t)(b_lo + kk * kstep);
              if (VARIANT == 2) {
                // A (128 rows x 16 bf16 per k-step = 8 columns of 32 bit) resident in tensor memory behind the two accumulators:
                // the K = 384 weight block of the 384-wide GEMMs would occupy 192 columns
                tc_mma_ts(d, tmem_base + 2u * N + (uint32_t)((kb * 4 + kk) % ((512 - 2 * N) / 8 > 0 ? (512 - 2 * N) / 8 : 1)) * 8u, bd, idesc, (kb > 0 || kk > 0) ? 1u : 0u);  // stay inside the 512 columns
              } else {
                tc_mma(d, ad, bd, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
              }
            }
            tc_commit(bars + stage * 8);
          }
          __syncwarp();
          if (++stage == STAGES) stage = 0;
        }
      }
      if (leader) tc_commit(final_bar);
      __syncwarp();
    }
    long long t1 = clock64();
    mbar_wait(final_bar, 0);
    long long t2 = clock64();
    if (lane == 0) {
      out[blockIdx.x * 2 + 0] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    if (lane == 0) asm volatile("st.volatile.shared.b32 [%0], %1;" ::"r"(bars + 112), "r"(1));
  } else if (warp >= 4) {
    // "epilogue" warps: drain accumulator columns with tcgen05.ld as fast as they can until the MMA warp is done.
    // ld_mode 1: read the accumulator stage the MMAs are NOT writing to most of the time (columns 256..383); 2: same columns
    if (ld_mode > 0) {
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (ld_mode == 1 ? 256u : 0u) + (warp >= 8 ? 64u : 0u);
      uint32_t done = 0, sink = 0;
      long long rounds = 0;
      while (!done) {
        uint32_t v[32], w[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31]) : "r"(taddr + 32u) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) sink ^= v[i] ^ w[i];
        ++rounds;
        asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(done) : "r"(bars + 112));
      }
      if (sink == 0x12345678u) out[0] = 0;
      if (lane == 0 && warp == 4) out[300 + blockIdx.x] = rounds;
    }
  } else if (warp == 2 && lane == 0 && traffic_bytes > 0) {
    // background L2 -> smem stream, two 8 KB copies in flight, until roughly the byte budget is spent
    const uint32_t b0 = bars + (STAGES + 1) * 8, b1 = bars + (STAGES + 2) * 8;
    const uint8_t* src = gsrc + (size_t)blockIdx.x * (1u << 20);
    uint32_t ph = 0;
    for (int off = 0; off < traffic_bytes; off += 16384) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b0), "r"(8192) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(traffic_dst), "l"(src + (off & 0xFFFFF)), "r"(8192), "r"(b0) : "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b1), "r"(8192) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(traffic_dst + 8192), "l"(src + ((off + 8192) & 0xFFFFF)), "r"(8192), "r"(b1) : "memory");
      mbar_wait(b0, ph);
      mbar_wait(b1, ph);
      ph ^= 1u;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// variant 2: a real producer/consumer ring.  One thread streams `a_bytes + b_bytes` per k-block into the stage with
// cp.async.bulk (full barrier), the MMA thread consumes it (4 MMAs) and frees the slot with tcgen05.commit (empty barrier).
// src_mode 0: every CTA streams its own 4 MB window (HBM + L2 misses); 1: all CTAs read the same 256 KB (L2 hits).
// ---------------------------------------------------------------------------------------------
template <int N, int NST>
__global__ void __launch_bounds__(128, 1) ring_bench_kernel(long long* out, int kblocks, const uint8_t* gsrc, int a_bytes, int b_bytes, int src_mode, int timed, long long* trace, int mps, int lps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = 128 * 64 * 2, B_BYTES = N * 64 * 2, STAGE = A_BYTES + B_BYTES;
  const uint32_t bars = base + NST * STAGE;  // full[NST], empty[NST], final
  const uint32_t tmem_slot = bars + (2 * NST + 2) * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < NST * STAGE / 4; i += blockDim.x) asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i * 4), "r"(0x3c003c00u));
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2 * NST + 1; ++s) mbar_init(bars + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t final_bar = bars + 2 * NST * 8;
  if (warp == 2 && lane == 0) {
    const size_t window = src_mode == 0 ? (4u << 20) : (256u << 10);
    const uint8_t* src = gsrc + (src_mode == 0 ? (size_t)blockIdx.x * window : 0);
    int stage = 0; uint32_t phase = 0; size_t off = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      const bool tr = trace && blockIdx.x == 0 && kb < 96;
      if (tr) trace[kb * 8 + 4] = clock64();
      mbar_wait(bars + (NST + stage) * 8, phase ^ 1u);
      if (tr) trace[kb * 8 + 5] = clock64();
      const uint32_t full = bars + stage * 8, dst = base + stage * STAGE;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"((a_bytes + b_bytes) * lps) : "memory");
      for (int l = 0; l < lps; ++l) {
        if (a_bytes) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src + off), "r"(a_bytes), "r"(full) : "memory");
        off = (off + a_bytes) & (window / 2 - 1);
        if (b_bytes) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + A_BYTES), "l"(src + off), "r"(b_bytes), "r"(full) : "memory");
        off = (off + b_bytes) & (window / 2 - 1);
      }
      if (tr) trace[kb * 8 + 6] = clock64();
      if (++stage == NST) { stage = 0; phase ^= 1u; }
    }
  } else if (warp == 1) {
    long long t0 = clock64(), twait = 0;
    {
      const bool leader = elect_one();
      const uint32_t hi = (uint32_t)((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        const bool tr = trace && blockIdx.x == 0 && kb < 96 && leader;
        long long w0 = ((timed & 1) || tr) ? clock64() : 0;
        if (!(timed & 2)) mbar_wait(bars + stage * 8, phase);
        if (!(timed & 4)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        long long w1 = ((timed & 1) || tr) ? clock64() : 0;
        if (timed & 1) twait += w1 - w0;
        if (tr) { trace[kb * 8 + 0] = w0; trace[kb * 8 + 1] = w1; }
        const uint32_t a_lo = (((base + stage * STAGE) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t b_lo = (((base + stage * STAGE + A_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
        const uint32_t d = tmem_base + (uint32_t)(((kb / 6) & 1) * N);
        if (leader) {
          for (int g = 0; g < mps; g += 4) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc_mma(d, ((uint64_t)hi << 32) | (a_lo + kk * 2), ((uint64_t)hi << 32) | (b_lo + kk * 2), idesc, (kb % 6 > 0 || kk > 0 || g > 0) ? 1u : 0u);
          }
          if (tr) trace[kb * 8 + 2] = clock64();
          tc_commit(bars + (NST + stage) * 8);
          if (tr) trace[kb * 8 + 3] = clock64();
        }
        __syncwarp();
        if (++stage == NST) { stage = 0; phase ^= 1u; }
      }
      if (leader) tc_commit(final_bar);
    }
    __syncwarp();
    mbar_wait(final_bar, 0);
    long long t2 = clock64();
    if (lane == 0) { out[blockIdx.x * 2 + 0] = twait; out[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

template <int N, int NST>
static void run_ring(long long* d_out, const uint8_t* gsrc, int a_bytes, int b_bytes, int src_mode, int flags = 0, bool do_trace = false, int mps = 4, int lps = 1) {
  constexpr int smem = NST * (128 * 64 * 2 + N * 64 * 2) + 256 + 1024;
  auto k = ring_bench_kernel<N, NST>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int kblocks = 600;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<148, 128, smem>>>(d_out, kblocks, gsrc, a_bytes, b_bytes, src_mode, flags, do_trace ? d_out + 512 : nullptr, mps, lps);
  cudaEventRecord(e0);
  k<<<148, 128, smem>>>(d_out, kblocks, gsrc, a_bytes, b_bytes, src_mode, flags, do_trace ? d_out + 512 : nullptr, mps, lps);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("ring: CUDA error %s\n", cudaGetErrorString(e)); return; }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[296];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double wait = 0, total = 0;
  for (int i = 0; i < 148; ++i) { wait += h[2 * i]; total += h[2 * i + 1]; }
  if (do_trace) {
    long long tr[96 * 8];
    cudaMemcpy(tr, d_out + 512, sizeof(tr), cudaMemcpyDeviceToHost);
    const long long z = tr[0];
    printf("  kb | consumer: wait_begin wait_end mma_issued commit_done | producer: wait_begin wait_end issued   (cycles since first event)\n");
    for (int kb = 0; kb < 40; ++kb)
      printf("  %2d | %7lld %7lld %7lld %7lld | %7lld %7lld %7lld\n", kb, tr[kb * 8] - z, tr[kb * 8 + 1] - z, tr[kb * 8 + 2] - z, tr[kb * 8 + 3] - z, tr[kb * 8 + 4] - z,
             tr[kb * 8 + 5] - z, tr[kb * 8 + 6] - z);
  }
  const double cyc = total / 148 / kblocks;
  printf("[%2d MMA/stage, %d x loads: %.1f cyc/MMA] ", mps, lps, cyc / mps);
  printf("ring flags=%d N=%3d stages=%d load %5d+%5d B/kblock src=%s | %.0f cyc/kblock (MMA floor %d), wait %.0f, %.1f B/clk/SM, chip %.2f TB/s\n", flags, N, NST, a_bytes, b_bytes,
         src_mode ? "L2 " : "HBM", cyc, 2 * N, wait / 148 / kblocks, (a_bytes + b_bytes) / cyc, 148.0 * kblocks * (a_bytes + b_bytes) / (ms * 1e-3) / 1e12);
}

template <int N, int VARIANT, bool MN>
static void run(const char* name, long long* d_out, const uint8_t* gsrc, int traffic, int ld_mode = 0, int nthreads = 128) {
  constexpr int smem = STAGES * (128 * 64 * 2 + N * 64 * 2) + 128 + 16384 + 1024;
  auto k = mma_bench_kernel<N, VARIANT, MN>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int tiles = 64;
  for (int rep = 0; rep < 2; ++rep) k<<<148, nthreads, smem>>>(d_out, tiles, gsrc, traffic, ld_mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return; }
  long long h[296];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double issue = 0, total = 0;
  for (int i = 0; i < 148; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
  const double n_mma = (double)tiles * KB_PER_TILE * 4;
  if (ld_mode) { long long r; cudaMemcpy(&r, d_out + 300, 8, cudaMemcpyDeviceToHost); printf("  [ld warps %d, mode %d: %.1f rounds of 2x(32 lanes x 32 cols) per tile per warp] ", (nthreads - 128) / 32, ld_mode, (double)r / tiles); }
  printf("%-34s N=%3d traffic=%8d B | issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d)\n", name, N, traffic, issue / 148 / n_mma, total / 148 / n_mma, N / 2);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  long long* d_out;
  uint8_t* gsrc;
  cudaMalloc(&d_out, (512 + 96 * 8) * sizeof(long long));
  cudaMemset(d_out, 0, (512 + 96 * 8) * sizeof(long long));
  cudaMalloc(&gsrc, 640u << 20);
  cudaMemset(gsrc, 0, 640u << 20);
  run<128, 0, false>("v0 lane0-branch K-major", d_out, gsrc, 0);
  run<128, 1, false>("v1 elect K-major", d_out, gsrc, 0);
  run<192, 1, false>("v1 elect K-major", d_out, gsrc, 0);
  run<256, 1, false>("v1 elect K-major", d_out, gsrc, 0);
  run<128, 0, true>("v0 lane0-branch MN-major", d_out, gsrc, 0);
  run<128, 1, true>("v1 elect MN-major", d_out, gsrc, 0);
  run<256, 1, true>("v1 elect MN-major", d_out, gsrc, 0);
  // with a TMA-like producer stream: bytes per CTA ~ what a resident-weights tile loop would pull (98 KB per tile)
  run<128, 1, false>("v1 elect K-major +stream", d_out, gsrc, 64 * 98304);
  run<256, 1, false>("v1 elect K-major +stream", d_out, gsrc, 64 * 98304);
  run<128, 1, true>("v1 elect MN-major +stream", d_out, gsrc, 64 * 98304);
  if (getenv("TS_ONLY")) {  // A from tensor memory vs A from shared memory
    run<128, 1, false>("SS: A smem, B smem (K-major)", d_out, gsrc, 0);
    run<128, 2, false>("TS: A tmem, B smem (K-major)", d_out, gsrc, 0);
    run<192, 2, false>("TS: A tmem, B smem (K-major)", d_out, gsrc, 0);
    run<128, 2, false>("TS +stream", d_out, gsrc, 64 * 98304);
    run<128, 1, false>("SS +stream", d_out, gsrc, 64 * 98304);
    return 0;
  }
  if (getenv("MPS_ONLY")) {
    for (int mps = 4; mps <= 16; mps *= 2) {
      run_ring<128, 6>(d_out, gsrc, 0, 0, 1, 0, false, mps, 1);
      run_ring<128, 6>(d_out, gsrc, 16384, 0, 1, 0, false, mps, mps / 4);       // resident-weights fwd: 16 KB per 4 MMAs
      run_ring<128, 6>(d_out, gsrc, 16384, 16384, 1, 0, false, mps, mps / 4);   // streaming / wgrad: 32 KB per 4 MMAs
      run_ring<128, 6>(d_out, gsrc, 16384, 16384, 0, 0, false, mps, mps / 4);   // same from HBM
    }
    run_ring<256, 4>(d_out, gsrc, 16384, 32768, 1, 0, false, 8, 2);
    return 0;
  }
  if (getenv("LD_ONLY")) {
    run<128, 1, false>("v1 K-major", d_out, gsrc, 0, 0, 128);
    run<128, 1, false>("v1 K-major + 4 ld warps other", d_out, gsrc, 0, 1, 256);
    run<128, 1, false>("v1 K-major + 8 ld warps other", d_out, gsrc, 0, 1, 384);
    run<128, 1, false>("v1 K-major + 4 ld warps same", d_out, gsrc, 0, 2, 256);
    run<128, 1, false>("v1 K-major + 8 ld warps same", d_out, gsrc, 0, 2, 384);
    return 0;
  }
  if (getenv("TRACE_ONLY")) {
    run_ring<128, 6>(d_out, gsrc, 0, 0, 1, 0, true);
    run_ring<128, 6>(d_out, gsrc, 16384, 16384, 1, 0, true);
    return 0;
  }
  printf("flags: 1 = clock reads around the wait, 2 = skip the full-barrier wait, 4 = skip tcgen05.fence::after_thread_sync\n");
  run_ring<128, 6>(d_out, gsrc, 0, 0, 1, 0);
  run_ring<128, 6>(d_out, gsrc, 0, 0, 1, 1);
  run_ring<128, 6>(d_out, gsrc, 0, 0, 1, 4);
  run_ring<128, 6>(d_out, gsrc, 16384, 0, 1, 4);
  run_ring<128, 6>(d_out, gsrc, 16384, 16384, 1, 4);
  run_ring<128, 6>(d_out, gsrc, 16384, 16384, 0, 4);
  for (int src = 1; src >= 0; --src) {
    run_ring<128, 6>(d_out, gsrc, 0, 0, src);
    run_ring<128, 6>(d_out, gsrc, 16384, 0, src);
    run_ring<128, 6>(d_out, gsrc, 16384, 8192, src);
    run_ring<128, 6>(d_out, gsrc, 16384, 16384, src);
    run_ring<256, 4>(d_out, gsrc, 0, 0, src);
    run_ring<256, 4>(d_out, gsrc, 16384, 0, src);
    run_ring<256, 4>(d_out, gsrc, 16384, 16384, src);
    run_ring<256, 4>(d_out, gsrc, 16384, 32768, src);
  }
  return 0;
}
