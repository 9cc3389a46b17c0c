// common.cuh — shared device/host helpers for libvitb200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vitb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvitb200 is written for sm_100a (B200) only"
#endif

namespace vitb {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define VITB_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::vitb::set_error(__VA_ARGS__);      \
      return -1;                           \
    }                                      \
  } while (0)

#define VITB_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::vitb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                   \
    }                                                                                   \
  } while (0)

// after a kernel launch (also counts launches: bench.py reports how many of OUR kernels ran per step)
void count_launch();
#define VITB_LAUNCH_OK()                   \
  do {                                     \
    ::vitb::count_launch();                \
    VITB_CUDA_OK(cudaPeekAtLastError());   \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (on by default; VITB_PDL=0 switches it off).  Every kernel calls pdl_wait() before its first access
// to memory another kernel may have written or may still read (griddepcontrol.wait returns when the preceding grid has completed
// and its memory is visible), so it is safe to launch it with the programmatic-stream-serialization attribute, which lets block
// scheduling, barrier initialisation, TMEM allocation and tensor-map prefetch overlap the predecessor's tail.  There is no early
// trigger (griddepcontrol.launch_dependents at kernel entry was measured 3 % SLOWER in round 1: early-resident dependents get in
// the way of the multi-wave attention kernels).  Measured in round 2 on the captured training graph (profiles/r2_pdl_ab.md):
// batch 1024 neutral (6.04 vs 6.05 ms/step), batch 128 1.609 -> 1.481 ms (+8.6 %), T = 17 / batch 1024 2.448 -> 2.280 ms (+7 %):
// the shorter the kernels, the more the launch-to-first-instruction latency matters.  Both instructions are no-ops in a plain launch.
// ---------------------------------------------------------------------------------------------
// No early trigger (griddepcontrol.launch_dependents): measured twice.  Round 1: at kernel entry, every kernel, 3 % slower at
// batch 1024.  Round 2: at entry only for grids that are resident at once (<= 296 CTAs, so that later waves never compete with
// early-resident dependents): 6.25 vs 5.97 ms at batch 1024, 1.60 vs 1.48 ms at batch 128, 2.57 vs 2.27 ms at T = 17
// (profiles/r2_pdl_ab.md) — dependents that become resident early sit in griddepcontrol.wait holding registers and shared
// memory the running kernel's neighbours could use.  -DVITB_PDL_EARLY_CTAS=296 rebuilds that experiment.
#ifndef VITB_PDL_EARLY_CTAS
#define VITB_PDL_EARLY_CTAS 0
#endif
__device__ __forceinline__ void pdl_trigger() {
  if (VITB_PDL_EARLY_CTAS > 0 && gridDim.x * gridDim.y * gridDim.z <= (unsigned)VITB_PDL_EARLY_CTAS) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();  // VITB_PDL=1 in the environment switches the launch attribute on

// Optional L2 persistence window (vitb_set_l2_persisting_window): every kernel launched while it is set carries it as a launch
// attribute (so it is recorded in captured graph nodes): reads inside the window are kept in the persisting carve-out of L2.
struct PersistWindow { void* base; size_t bytes; float hit_ratio; };
const PersistWindow& persist_window();
void set_persist_window(void* base, size_t bytes, float hit_ratio);
inline int add_persist_attr(cudaLaunchAttribute* attr, int n) {
  const PersistWindow& w = persist_window();
  if (w.base == nullptr || w.bytes == 0) return n;
  attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
  attr[n].val.accessPolicyWindow.base_ptr = w.base;
  attr[n].val.accessPolicyWindow.num_bytes = w.bytes;
  attr[n].val.accessPolicyWindow.hitRatio = w.hit_ratio;
  attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  return n + 1;
}

template <typename... Exp, typename... Act>
inline cudaError_t launch_kernel(void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Act&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  n = add_persist_attr(attr, n);
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Exp>(args)...);
}
// same, as thread-block clusters of `cluster` consecutive CTAs (cta_group::2 GEMM pairs)
template <typename... Exp, typename... Act>
inline cudaError_t launch_kernel_cluster(void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster, Act&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Exp>(args)...);
}
// kernel expression in parentheses (template arguments contain commas)
#define VITB_LAUNCH(kernel, grid, block, smem, stream, ...) (void)::vitb::launch_kernel(kernel, grid, block, smem, stream, __VA_ARGS__)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// activation storage types
// ---------------------------------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

template <typename T> struct Act;
template <> struct Act<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct Act<bf16> {
  static __device__ __forceinline__ float ld(const bf16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// 4 consecutive activations <-> float4 (8-byte access for bf16, 16-byte for fp32)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 a = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&a);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}

// ---------------------------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------------------------
// exact (erf) GELU, nn.GELU() default (layers.py:34,37), evaluated without erff():
//   Phi(-u) = 0.5*erfc(u/sqrt2) = exp2(r(u)),  r = degree-6 polynomial fitted on u in [0,10]  (max abs error of
//   Phi 2.0e-7, of gelu 1.8e-7, of gelu' 2.0e-7 over R: below fp32 rounding of the surrounding arithmetic)
//   Phi(x) = x >= 0 ? 1 - Phi(-x) : Phi(-|x|).     6 FMA + 1 MUFU.EX2 instead of erff's two polynomial branches.
// 2^x as ONE MUFU.EX2 (ex2.approx.ftz: relative error 2^-22, denormal results flush to 0).  exp2f() wraps the same instruction in
// range guards (compare, scale by 0.5, square) that triple the instruction count of the GELU epilogues, which are ALU-bound.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float normal_cdf_f(float x) {
  const float u = fminf(fabsf(x), 10.0f);
  float r = 2.641155697e-05f;
  r = fmaf(r, u, -7.098118658e-04f);
  r = fmaf(r, u, 7.883246057e-03f);
  r = fmaf(r, u, -5.310586467e-02f);
  r = fmaf(r, u, -4.589958787e-01f);
  r = fmaf(r, u, -1.151131988e+00f);
  r = fmaf(r, u, -9.999994636e-01f);
  const float e = ex2_approx(r);
  return x >= 0.0f ? 1.0f - e : e;
}
__device__ __forceinline__ float gelu_f(float x) { return x * normal_cdf_f(x); }
// d/dx [x Phi(x)] = Phi(x) + x phi(x),  phi(x) = exp2(-x^2 * log2(e)/2) / sqrt(2 pi)
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float pdf = 0.39894228040143267794f * ex2_approx(-0.72134752044448170368f * x * x);
  return fmaf(x, pdf, normal_cdf_f(x));
}

// The same two functions for the tensor-core epilogues, whose results are rounded to bf16 (relative 2^-9) and which are bound by
// instruction issue (8 epilogue warps x 64 elements per lane per tile): a degree-4 exponent polynomial on u = min(|x|, 5.5)
// (max abs error of Phi 2.9e-5, of gelu 1.2e-5, relative error of gelu below 2.5e-5 for x > 0.05 — 150x under the bf16 rounding of
// the stored value) and  gelu(x) = max(x, 0) - u Phi(-u)  instead of a select: 8 instructions per element instead of 12 (14 / 16
// for the derivative).  The fp32 check mode and the stand-alone GELU-backward kernel keep the accurate forms above.
__device__ __forceinline__ float phi_neg_bf16_f(float u) {  // Phi(-u), u in [0, 5.5]
  float r = 4.311318975e-03f;
  r = fmaf(r, u, -4.634770751e-02f);
  r = fmaf(r, u, -4.642629325e-01f);
  r = fmaf(r, u, -1.149653077e+00f);
  r = fmaf(r, u, -1.000083923e+00f);
  return ex2_approx(r);
}
__device__ __forceinline__ float gelu_bf16_f(float x) {
  const float u = fminf(fabsf(x), 5.5f);
  return fmaf(-u, phi_neg_bf16_f(u), fmaxf(x, 0.0f));
}
__device__ __forceinline__ float gelu_grad_bf16_f(float x) {
  const float u = fminf(fabsf(x), 5.5f);
  const float e = phi_neg_bf16_f(u);
  const float pdf = ex2_approx(-0.72134752044448170368f * (x * x));
  const float cdf = x >= 0.0f ? 1.0f - e : e;
  return fmaf(0.39894228040143267794f * x, pdf, cdf);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// fixed-order reduction of per-block partials: out_k[c] = sum_p ws[k][p][c]   (k = blockIdx.y < 3)
// (template so that every translation unit can instantiate it without relocatable device code)
// ---------------------------------------------------------------------------------------------
// block (32, 8): x = 4-column lane, y = slice of the partials.  `bx` of `nbx` blocks cover `cols` columns.
__device__ __forceinline__ void finalize_block_cols(const float* __restrict__ base, int nparts, int64_t cols, float* __restrict__ out, int bx, int nbx) {
  __shared__ float4 red[8][32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  if ((cols & 3) == 0) {
    const int64_t c = ((int64_t)bx * 32 + tx) * 4;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    if (c < cols) {
      const float* p = base + c;
      int i = ty;
      for (; i + 8 < nparts; i += 16) {
        const float4 a0 = *reinterpret_cast<const float4*>(p + (size_t)i * cols);
        const float4 a1 = *reinterpret_cast<const float4*>(p + (size_t)(i + 8) * cols);
        s0.x += a0.x; s0.y += a0.y; s0.z += a0.z; s0.w += a0.w;
        s1.x += a1.x; s1.y += a1.y; s1.z += a1.z; s1.w += a1.w;
      }
      if (i < nparts) {
        const float4 a0 = *reinterpret_cast<const float4*>(p + (size_t)i * cols);
        s0.x += a0.x; s0.y += a0.y; s0.z += a0.z; s0.w += a0.w;
      }
    }
    red[ty][tx] = make_float4(s0.x + s1.x, s0.y + s1.y, s0.z + s1.z, s0.w + s1.w);
    __syncthreads();
    if (ty == 0 && c < cols) {
      float4 r = red[0][tx];
#pragma unroll
      for (int j = 1; j < 8; ++j) {
        const float4 o = red[j][tx];
        r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
      }
      *reinterpret_cast<float4*>(out + c) = r;
    }
    return;
  }
  // generic (cols not a multiple of 4): one thread per column, strided over the blocks
  const int tid = ty * 32 + tx;
  for (int64_t c = (int64_t)bx * 256 + tid; c < cols; c += (int64_t)nbx * 256) {
    float s = 0.f;
    for (int i = 0; i < nparts; ++i) s += base[(size_t)i * cols + c];
    out[c] = s;
  }
}

// launch with block (32, 8) and grid (ceil(cols / 128), number of outputs)
template <int kUnused = 0>
__global__ void __launch_bounds__(256) partials_finalize_kernel(const float* __restrict__ ws, int nparts, int64_t cols,
                                                                float* out0, float* out1, float* out2) {
  pdl_trigger();
  pdl_wait();
  const int k = blockIdx.y;
  float* out = k == 0 ? out0 : (k == 1 ? out1 : out2);
  if (out == nullptr) return;
  finalize_block_cols(ws + (size_t)k * nparts * cols, nparts, cols, out, blockIdx.x, gridDim.x);
}

// two partial sets with different widths in one launch (wgrad: dW and the bias gradient): grid = blocks(cols0) + blocks(cols1)
template <int kUnused = 0>
__global__ void __launch_bounds__(256) partials_finalize2_kernel(const float* __restrict__ ws0, int64_t cols0, float* out0, const float* __restrict__ ws1,
                                                                 int64_t cols1, float* out1, int nparts) {
  pdl_trigger();
  pdl_wait();
  const int nb0 = (int)((cols0 + 127) / 128);
  if ((int)blockIdx.x < nb0) finalize_block_cols(ws0, nparts, cols0, out0, blockIdx.x, nb0);
  else finalize_block_cols(ws1, nparts, cols1, out1, blockIdx.x - nb0, gridDim.x - nb0);
}

// "tall" variant for MANY partial rows of FEW columns (LayerNorm / GELU column sums: 300-740 partials x 384): block (8, 64) =
// 8 four-column lanes x 64 slices of the partials, ceil(cols / 32) blocks per output.  The plain kernel would run such a shape on 3
// blocks whose threads each walk ~90 dependent L2 round trips (21 us for 1 MB); here a thread walks nparts / 64 of them.
// Slices are combined in a fixed order (8 groups of 8, then the 8 group sums): deterministic.
__device__ __forceinline__ void finalize_tall_block(const float* __restrict__ base, int nparts, int64_t cols, float* __restrict__ out, int bx) {
  __shared__ float4 red[64][8];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t c = ((int64_t)bx * 8 + tx) * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cols) {
    for (int i = ty; i < nparts; i += 64) {
      const float4 a = *reinterpret_cast<const float4*>(base + (size_t)i * cols + c);
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty < 8) {  // group ty sums slices 8 ty .. 8 ty + 7
    float4 r = red[8 * ty][tx];
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      const float4 o = red[8 * ty + j][tx];
      r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
    }
    red[8 * ty][tx] = r;
  }
  __syncthreads();
  if (ty == 0 && c < cols) {
    float4 r = red[0][tx];
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      const float4 o = red[8 * j][tx];
      r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
    }
    *reinterpret_cast<float4*>(out + c) = r;
  }
}

template <int kUnused = 0>
__global__ void __launch_bounds__(512) partials_finalize_tall_kernel(const float* __restrict__ ws, int nparts, int64_t cols, float* out0, float* out1,
                                                                     float* out2) {
  pdl_trigger();
  pdl_wait();
  const int k = blockIdx.y;
  float* out = k == 0 ? out0 : (k == 1 ? out1 : out2);
  if (out == nullptr) return;
  finalize_tall_block(ws + (size_t)k * nparts * cols, nparts, cols, out, blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// Deferred reductions (vitb_defer_begin / vitb_defer_flush): between the two calls every split reduction whose result is only a
// GRADIENT (wgrad dW / db partials, LayerNorm dgamma / dbeta / column sums, GELU-backward column sums) places its fp32 partials in
// the caller's arena and records a job instead of launching its second pass; the flush reduces all recorded jobs with the same
// per-column summation order in one launch per job class — ~45 launches and their dependent-launch latencies less per step.
// ---------------------------------------------------------------------------------------------
struct ReduceJob {
  const float* src;  // [nparts][cols] fp32 partials
  float* dst;        // [cols]
  int nparts;
  int cols;
};
constexpr int kMaxReduceJobs = 120;  // per launch: the table travels in the kernel parameters (< 4 KB)
struct ReduceBatch {
  int njobs;
  int block_start[kMaxReduceJobs + 1];  // first block of every job
  ReduceJob jobs[kMaxReduceJobs];
};
bool defer_active();
void* defer_alloc(size_t bytes);          // arena memory for partials (256-byte aligned), or nullptr (not deferring / arena exhausted)
bool defer_owns(const void* p);
void defer_add(const float* src, int nparts, int64_t cols, float* dst);
inline bool finalize_is_tall(int nparts, int64_t cols) { return (cols & 3) == 0 && nparts >= 64 && cols <= 8192; }

inline dim3 finalize_grid(int64_t cols, int nout) { return dim3((unsigned)((cols + 127) / 128), (unsigned)nout); }
// picks the tall kernel for many partials of a narrow output, the plain one otherwise (one launch either way); partials that live
// in the deferral arena are recorded for vitb_defer_flush instead.  Counts its own launch.
inline cudaError_t launch_finalize(const float* ws, int nparts, int64_t cols, float* out0, float* out1, float* out2, int nout, cudaStream_t st) {
  if (defer_active() && defer_owns(ws)) {
    float* outs[3] = {out0, out1, out2};
    for (int k = 0; k < nout; ++k)
      if (outs[k] != nullptr) defer_add(ws + (size_t)k * nparts * cols, nparts, cols, outs[k]);
    return cudaSuccess;
  }
  count_launch();
  if (finalize_is_tall(nparts, cols))
    return launch_kernel(partials_finalize_tall_kernel<0>, dim3((unsigned)((cols + 31) / 32), (unsigned)nout), dim3(8, 64), 0, st, ws, nparts, cols, out0,
                         out1, out2);
  return launch_kernel(partials_finalize_kernel<0>, finalize_grid(cols, nout), dim3(32, 8), 0, st, ws, nparts, cols, out0, out1, out2);
}
// dW and the bias gradient of a wgrad (two partial sets of different widths): one launch, or two deferred jobs
inline cudaError_t launch_finalize2(const float* ws0, int64_t cols0, float* out0, const float* ws1, int64_t cols1, float* out1, int nparts, cudaStream_t st) {
  if (defer_active() && defer_owns(ws0) && defer_owns(ws1)) {
    defer_add(ws0, nparts, cols0, out0);
    defer_add(ws1, nparts, cols1, out1);
    return cudaSuccess;
  }
  count_launch();
  const unsigned nb = (unsigned)((cols0 + 127) / 128 + (cols1 + 127) / 128);
  return launch_kernel(partials_finalize2_kernel<0>, dim3(nb), dim3(32, 8), 0, st, ws0, cols0, out0, ws1, cols1, out1, nparts);
}
inline dim3 finalize_block() { return dim3(32, 8); }

// ---------------------------------------------------------------------------------------------
// Adam (coupled L2), torch.optim.Adam single-tensor arithmetic: shared by the flat-buffer kernel (loss_adam.cu) and the
// data-parallel reduce + Adam + broadcast kernel (dp.cu)
// ---------------------------------------------------------------------------------------------
struct AdamHyper {
  float step_size, bc2_sqrt, beta1, beta2, eps, wd, grad_scale, one_minus_beta1, one_minus_beta2;
};
__host__ __device__ __forceinline__ AdamHyper adam_hyper_from(const float* h9) {
  AdamHyper h;
  h.step_size = h9[0]; h.bc2_sqrt = h9[1]; h.beta1 = h9[2]; h.beta2 = h9[3]; h.eps = h9[4]; h.wd = h9[5]; h.grad_scale = h9[6];
  h.one_minus_beta1 = h9[7]; h.one_minus_beta2 = h9[8];
  return h;
}
// Explicit rounding intrinsics: the compiler may not re-associate or contract differently in different kernels, so every kernel
// that inlines this function produces the same bits for the same inputs.
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamHyper& h) {
  g = __fmaf_rn(g, h.grad_scale, __fmul_rn(h.wd, p));
  m = __fmaf_rn(__fsub_rn(g, m), h.one_minus_beta1, m);                             // exp_avg.lerp_(grad, 1 - beta1)
  v = __fmaf_rn(__fmul_rn(h.one_minus_beta2, g), g, __fmul_rn(v, h.beta2));         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), h.bc2_sqrt), h.eps);
  p = __fmaf_rn(-h.step_size, __fdiv_rn(m, denom), p);
}

// torch.optim.SGD with momentum (network.py:78-84: momentum = beta1, dampening 0, no Nesterov) on the same hyper block:
// step_size = lr, beta1 = momentum, wd, grad_scale.  buf starts at zero, so the first step gives buf = g like torch's clone.
__device__ __forceinline__ void sgd_one(float& p, float g, float& buf, const AdamHyper& h) {
  g = __fmaf_rn(g, h.grad_scale, __fmul_rn(h.wd, p));
  buf = __fmaf_rn(h.beta1, buf, g);
  p = __fmaf_rn(-h.step_size, buf, p);
}

// ---------------------------------------------------------------------------------------------
// nn.Dropout keep masks (layers.py:35, 38, 102), shared by the stand-alone pass (elementwise.cu) and the fused sites (GEMM
// epilogues, GELU backward, LayerNorm backward).  The mask is never stored: it is a pure function of (seed, site, step, element
// index) — Philox4x32-10 with key = seed and counter = (group lo, group hi, site, step), one call per group of eight consecutive
// elements of the flattened (rows, cols) tensor, sixteen bits per element (element j of the group uses bits 16 (j & 1) .. of word
// j >> 1; kept iff that value >= round(p * 65536)).  `step` comes from device memory when step_dev != NULL, so a captured CUDA
// graph draws a fresh mask on every replay.
// ---------------------------------------------------------------------------------------------
struct DropParams {
  uint32_t thr;   // 0 = no dropout at this site
  float scale;    // 1 / (1 - p)
  uint2 key;
  uint32_t site, step;
  const uint32_t* step_dev;
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ uint32_t drop_step(const DropParams& d) { return d.step_dev != nullptr ? *d.step_dev : d.step; }
// the four mask words of element group g (elements 8 g .. 8 g + 7)
__device__ __forceinline__ uint4 drop_words(const DropParams& d, uint32_t step, uint64_t g) {
  return philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), d.site, step), d.key);
}
// element j (0..7) of the group whose words are w
__device__ __forceinline__ float drop_one(float v, uint32_t word, int j, uint32_t thr, float scale) {
  return ((word >> (16 * (j & 1))) & 0xffffu) >= thr ? v * scale : 0.f;
}
__device__ __forceinline__ void drop_apply8(float* f, const uint4& r, uint32_t thr, float scale) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = drop_one(f[j], w[j >> 1], j, thr, scale);
}
// four consecutive elements starting at flat element index e (a multiple of 4): half a mask group
__device__ __forceinline__ void drop_apply4(float4& v, const DropParams& d, uint32_t step, uint64_t e) {
  const uint4 r = drop_words(d, step, e >> 3);
  const bool hi = ((e >> 2) & 1) != 0;
  const uint32_t w0 = hi ? r.z : r.x, w1 = hi ? r.w : r.y;
  v.x = drop_one(v.x, w0, 0, d.thr, d.scale);
  v.y = drop_one(v.y, w0, 1, d.thr, d.scale);
  v.z = drop_one(v.z, w1, 0, d.thr, d.scale);
  v.w = drop_one(v.w, w1, 1, d.thr, d.scale);
}
// host: the kernel-side description of a vitb_dropout_t (NULL or p == 0: off).  Returns false on a bad p.
static inline bool make_drop_params(const vitb_dropout_t* d, DropParams* out) {
  *out = DropParams{};
  if (d == nullptr || d->p == 0.f) return true;
  if (!(d->p > 0.f && d->p < 1.f)) return false;
  out->thr = vitb_dropout_threshold(d->p);
  out->scale = 1.0f / (1.0f - d->p);
  out->key = make_uint2((uint32_t)d->seed, (uint32_t)(d->seed >> 32));
  out->site = d->site;
  out->step = d->step;
  out->step_dev = d->step_dev;
  return true;
}

// ---------------------------------------------------------------------------------------------
// GEMM epilogue description shared by the SIMT (fp32 check mode / small shapes) and tcgen05 paths
// ---------------------------------------------------------------------------------------------
enum EpiMode { EPI_FWD = 0, EPI_DGRAD = 1, EPI_RAW_F32 = 2 };

struct EpiParams {
  int mode;              // EpiMode
  int gelu;              // EPI_FWD: apply exact GELU after bias
  int out_f32;           // EPI_FWD: `out` is fp32 regardless of the activation type
  const float* bias;     // [N] or null
  const void* residual;  // [M,N] act or null (added after GELU)
  void* out;             // [M,N] act (or fp32 for RAW/out_f32)
  void* preact;          // [M,N] act or null: value before GELU
  const void* aux;       // EPI_DGRAD: z [M,N] act or null -> multiply by gelu'(z)
  int64_t ldc;           // row pitch of out/residual/preact/aux in elements
  // output row remap (patch embedding writes token rows of a (B,T,H) tensor and adds pos_emb):
  // phys_row = (m / rm_group) * rm_stride + rm_offset + m % rm_group ; pos row = phys_row % rm_stride
  int rm_group, rm_stride, rm_offset;
  const float* pos;      // [rm_stride, N] fp32 or null
};

}  // namespace vitb
