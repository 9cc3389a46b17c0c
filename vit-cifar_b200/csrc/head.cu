// head.cu — the classifier head `fc[1] = nn.Linear(hidden, num_classes)` (vit.py:63, 76) and its backward.
//
// N = num_classes is 10 or 100: far too narrow for a 128-wide tensor-core tile and, as a 64x64 SIMT tile, a 16-CTA
// latency-bound launch.  These kernels are plain FFMA with fixed-order reductions (bit-reproducible, used by both the bf16
// path and the fp32 check mode), shaped so that every SM has work:
//   fwd    logits[b, c] = bias[c] + sum_k a[b, k] w[c, k]          one warp per image
//   dgrad  da[b, k]     = sum_c dlogits[b, c] w[c, k]              one thread per (image, 4 columns)
//   wgrad  dw[c, k]     = sum_b dlogits[b, c] a[b, k],  dbias[c] = sum_b dlogits[b, c]
//                                                                  32 columns x 32 batch slices per CTA, slices combined in order
#include "common.cuh"
#include "gemm_internal.h"

namespace vitb {

template <typename T>
__global__ void __launch_bounds__(256) head_fwd_kernel(const T* __restrict__ a, const T* __restrict__ w, const float* __restrict__ bias,
                                                       float* __restrict__ out, int M, int N, int K) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const T* ar = a + (size_t)row * K;
  for (int c = 0; c < N; ++c) {
    const T* wr = w + (size_t)c * K;
    float acc = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 x = ld4(ar + k), y = ld4(wr + k);
      acc = fmaf(x.x, y.x, acc);
      acc = fmaf(x.y, y.y, acc);
      acc = fmaf(x.z, y.z, acc);
      acc = fmaf(x.w, y.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[(size_t)row * N + c] = acc + (bias ? bias[c] : 0.f);
  }
}

// K = 128 KV: a warp takes TWO rows, keeps them in registers and walks the classes four at a time — eight independent dot
// products and eight interleaved warp reductions in flight instead of one (the one-row, one-class-at-a-time kernel above is a chain
// of N dependent load -> FMA -> shuffle rounds: 104 us for the 100-class head at B = 1024, 4.7 % of that configuration's step).
// Per (row, class) the summation order is the one of head_fwd_kernel (lane-strided k, then the xor tree): bit-identical results.
constexpr int kHeadFwdClasses = 20;  // classes per CTA (a multiple of 4)
template <typename T, int KV>
__global__ void __launch_bounds__(256) head_fwd_rows_kernel(const T* __restrict__ a, const T* __restrict__ w, const float* __restrict__ bias,
                                                            float* __restrict__ out, int M, int N) {
  pdl_trigger();
  pdl_wait();
  constexpr int K = 128 * KV;
  const int lane = threadIdx.x & 31;
  const int r0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2;
  if (r0 >= M) return;
  const bool two = r0 + 1 < M;
  float4 x0[KV], x1[KV];
#pragma unroll
  for (int j = 0; j < KV; ++j) {
    x0[j] = ld4(a + (size_t)r0 * K + (j * 32 + lane) * 4);
    x1[j] = ld4(a + (size_t)(two ? r0 + 1 : r0) * K + (j * 32 + lane) * 4);
  }
  // blockIdx.y owns kHeadFwdClasses classes: the class loop is a chain of load -> FMA -> shuffle rounds (about a microsecond per
  // four classes), so a 100-class head is spread over five times as many warps instead of walked by one
  const int c_end = min(N, ((int)blockIdx.y + 1) * kHeadFwdClasses);
  for (int c = (int)blockIdx.y * kHeadFwdClasses; c < c_end; c += 4) {
    float s0[4], s1[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const T* wr = w + (size_t)min(c + cc, N - 1) * K;  // (classes beyond N: computed on the last row, not stored)
      float u = 0.f, v = 0.f;
#pragma unroll
      for (int j = 0; j < KV; ++j) {
        const float4 y = ld4(wr + (j * 32 + lane) * 4);
        u = fmaf(x0[j].x, y.x, u); u = fmaf(x0[j].y, y.y, u); u = fmaf(x0[j].z, y.z, u); u = fmaf(x0[j].w, y.w, u);
        v = fmaf(x1[j].x, y.x, v); v = fmaf(x1[j].y, y.y, v); v = fmaf(x1[j].z, y.z, v); v = fmaf(x1[j].w, y.w, v);
      }
      s0[cc] = u;
      s1[cc] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        s0[cc] += __shfl_xor_sync(0xffffffffu, s0[cc], o);
        s1[cc] += __shfl_xor_sync(0xffffffffu, s1[cc], o);
      }
    }
    if (lane < 4 && c + lane < c_end) {
      const float b = bias ? bias[c + lane] : 0.f;
      const float t0 = lane == 0 ? s0[0] : lane == 1 ? s0[1] : lane == 2 ? s0[2] : s0[3];
      const float t1 = lane == 0 ? s1[0] : lane == 1 ? s1[1] : lane == 2 ? s1[2] : s1[3];
      out[(size_t)r0 * N + c + lane] = t0 + b;
      if (two) out[(size_t)(r0 + 1) * N + c + lane] = t1 + b;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) head_dgrad_kernel(const float* __restrict__ dy, const T* __restrict__ w, T* __restrict__ dx, int M, int N,
                                                         int K) {
  pdl_trigger();
  pdl_wait();
  const int k4 = K >> 2;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * k4) return;
  const int row = (int)(i / k4), k = (int)(i % k4) * 4;
  const float* d = dy + (size_t)row * N;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < N; ++c) {
    const float g = d[c];
    const float4 y = ld4(w + (size_t)c * K + k);
    acc.x = fmaf(g, y.x, acc.x);
    acc.y = fmaf(g, y.y, acc.y);
    acc.z = fmaf(g, y.z, acc.z);
    acc.w = fmaf(g, y.w, acc.w);
  }
  st4(dx + (size_t)row * K + k, acc);
}

constexpr int kHeadCC = 8;       // classes per register pass
constexpr int kHeadSlices = 32;  // batch slices per CTA (the loop over the batch is latency-bound: 32 images per thread at B = 1024)

// grid (K / 32, ceil(N / kHeadCC)): a CTA owns 32 columns of x and kHeadCC classes, so that a 100-class head spreads over
// 12 x 13 CTAs instead of walking 13 class passes (each re-reading x) on 12 (162 -> ~10 us at B = 1024, C = 100)
template <typename T>
__global__ void __launch_bounds__(32 * kHeadSlices) head_wgrad_kernel(const float* __restrict__ dy, const T* __restrict__ x, float* __restrict__ dw,
                                                                      float* __restrict__ dbias, int M, int N, int K) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[kHeadSlices][kHeadCC][32];
  const int lane = threadIdx.x & 31, s = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  const int c0 = blockIdx.y * kHeadCC;
  {
    float acc[kHeadCC];
#pragma unroll
    for (int j = 0; j < kHeadCC; ++j) acc[j] = 0.f;
    if (k < K) {
#pragma unroll 4  // four images' loads in flight per thread: the loop is latency-bound
      for (int b = s; b < M; b += kHeadSlices) {
        const float xv = Act<T>::ld(x + (size_t)b * K + k);
        const float* d = dy + (size_t)b * N + c0;
#pragma unroll
        for (int j = 0; j < kHeadCC; ++j)
          if (c0 + j < N) acc[j] = fmaf(d[j], xv, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < kHeadCC; ++j) red[s][j][lane] = acc[j];
    __syncthreads();
    // thread (s, lane) with s < kHeadCC combines class s of this CTA for column `lane`, slices in fixed order
    for (int j = s; j < kHeadCC; j += kHeadSlices) {
      if (c0 + j < N && k < K) {
        float t = red[0][j][lane];
#pragma unroll
        for (int q = 1; q < kHeadSlices; ++q) t += red[q][j][lane];
        dw[(size_t)(c0 + j) * K + k] = t;
      }
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && dbias != nullptr) {
    // dbias of this CTA's classes: lane = class, slice = batch slice, slices combined in fixed order
    float* r2 = &red[0][0][0];  // [kHeadSlices][32]
    const int c = c0 + lane;
    float t = 0.f;
    if (lane < kHeadCC && c < N) {
#pragma unroll 4
      for (int b = s; b < M; b += kHeadSlices) t += dy[(size_t)b * N + c];
    }
    r2[s * 32 + lane] = t;
    __syncthreads();
    if (s == 0 && lane < kHeadCC && c < N) {
      float u = r2[lane];
#pragma unroll
      for (int q = 1; q < kHeadSlices; ++q) u += r2[q * 32 + lane];
      dbias[c] = u;
    }
  }
}

bool head_shape_ok(int M, int N, int K) { return N <= 128 && K % 4 == 0 && M > 0; }

template <typename T>
static int head_fwd_launch_t(const T* a, const T* w, const float* bias, float* out, int M, int N, int K, cudaStream_t st) {
  const int wpb = 8;
#define VITB_HEAD_ROWS(KV)                                                                                             \
  case KV: VITB_LAUNCH((head_fwd_rows_kernel<T, KV>), dim3(ceil_div(M, 2 * wpb), ceil_div(N, kHeadFwdClasses)), 32 * wpb, 0, st, a, w, bias, out, M, N); break;
  if (K % 128 == 0 && ((uintptr_t)a | (uintptr_t)w) % 16 == 0) {
    switch (K / 128) {
      VITB_HEAD_ROWS(1) VITB_HEAD_ROWS(2) VITB_HEAD_ROWS(3) VITB_HEAD_ROWS(4) VITB_HEAD_ROWS(6) VITB_HEAD_ROWS(8)
      default: VITB_LAUNCH((head_fwd_kernel<T>), ceil_div(M, wpb), 32 * wpb, 0, st, a, w, bias, out, M, N, K); break;
    }
  } else {
    VITB_LAUNCH((head_fwd_kernel<T>), ceil_div(M, wpb), 32 * wpb, 0, st, a, w, bias, out, M, N, K);
  }
#undef VITB_HEAD_ROWS
  VITB_LAUNCH_OK();
  return 0;
}

int head_fwd_launch(const void* a, const void* w, const float* bias, float* out, int M, int N, int K, int dt, cudaStream_t st) {
  if (dt == VITB_BF16) return head_fwd_launch_t<bf16>((const bf16*)a, (const bf16*)w, bias, out, M, N, K, st);
  return head_fwd_launch_t<float>((const float*)a, (const float*)w, bias, out, M, N, K, st);
}

int head_dgrad_launch(const float* dy, const void* w, void* dx, int M, int N, int K, int dt, cudaStream_t st) {
  const int64_t n = (int64_t)M * (K / 4);
  const int blocks = (int)ceil_div64(n, 256);
  if (dt == VITB_BF16) VITB_LAUNCH((head_dgrad_kernel<bf16>), blocks, 256, 0, st, dy, (const bf16*)w, (bf16*)dx, M, N, K);
  else VITB_LAUNCH((head_dgrad_kernel<float>), blocks, 256, 0, st, dy, (const float*)w, (float*)dx, M, N, K);
  VITB_LAUNCH_OK();
  return 0;
}

int head_wgrad_launch(const float* dy, const void* x, float* dw, float* dbias, int M, int N, int K, int dt, cudaStream_t st) {
  const dim3 blocks(ceil_div(K, 32), ceil_div(N, kHeadCC));
  if (dt == VITB_BF16) VITB_LAUNCH((head_wgrad_kernel<bf16>), blocks, 32 * kHeadSlices, 0, st, dy, (const bf16*)x, dw, dbias, M, N, K);
  else VITB_LAUNCH((head_wgrad_kernel<float>), blocks, 32 * kHeadSlices, 0, st, dy, (const float*)x, dw, dbias, M, N, K);
  VITB_LAUNCH_OK();
  return 0;
}

}  // namespace vitb
