// gemm_internal.h — host-side interfaces shared between gemm_simt.cu and gemm_tc.cu (not part of the C ABI).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace vitb {

// C[M,N] = sum_k A(m,k) * B(k,n), generic strides (elements).
//   A(m,k) = a[m*a_sm + kmap(k)],  kmap(k) = k*a_sk, or with a_kgroup > 0:
//            (k / a_kgroup) * a_kgroup_stride + (k % a_kgroup) * a_sk + a_koff
//   B(k,n) = b[k*b_sk + n*b_sn]
//   gather: operand is the patch matrix of an NCHW fp32 image (vit.py:79-89):
//            words(m, f) = img[b][c][ph*ps+kh][pw*ps+kw], m = (b*P+ph)*P+pw, f = (kh*ps+kw)*3+c
//            a_gather: A(m,k) = words(m,k);  b_gather: B(k,n) = words(k,n)
struct SimtGemmArgs {
  const void* a;
  const void* b;
  int M, N, K;
  int64_t a_sm, a_sk;
  int a_kgroup;
  int64_t a_kgroup_stride, a_koff;
  int64_t b_sk, b_sn;
  int a_gather, b_gather, gather_S, gather_P;
  EpiParams e;
};

int simt_gemm_launch(const SimtGemmArgs& g, int a_dt, int b_dt, int o_dt, int splits, cudaStream_t st);
int simt_pick_splits(int tiles, int K);
// 3-D bf16 tensor map {d0 (contiguous), d1, d2} with byte strides of d1 / d2, box {box0, box1, box2}, swizzle_bytes in {0, 32, 64, 128}
// (implemented in gemm_tc.cu, which owns the driver entry point)
int make_tma_map_3d_bf16(CUtensorMap* map, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes, uint64_t stride2_bytes,
                         uint32_t box0, uint32_t box1, uint32_t box2, int swizzle_bytes);

// classifier head (head.cu): N = num_classes <= 128, fp32 logits / dlogits, activations and weight in `dt`
bool head_shape_ok(int M, int N, int K);
int head_fwd_launch(const void* a, const void* w, const float* bias, float* out, int M, int N, int K, int dt, cudaStream_t st);
int head_dgrad_launch(const float* dy, const void* w, void* dx, int M, int N, int K, int dt, cudaStream_t st);
int head_wgrad_launch(const float* dy, const void* x, float* dw, float* dbias, int M, int N, int K, int dt, cudaStream_t st);
int colsum_small_launch(const void* x, float* out, int rows, int cols, int64_t ld, int x_dt, cudaStream_t st);

// tensor-core pieces of the patch embedding (gemm_tc.cu), bf16 only
//   fwd:  out[b, off + n, :] = words[b PP + n, :] · w[H, K]ᵀ + bias + pos[off + n, :]   (K = 48 / 192: partial k-block, TMA zero-fills;
//         stored straight into the (B, Tn, H) tensor through a 3-D tensor map)
//   bwd:  dw[H, K] = sum_m dout[b, off + n, :]ᵀ words[m, :] (A read through a 3-D tensor map that skips the cls rows),
//         dbias[H] = column sums of those rows (ones-operand MMA)
bool tc_patch_ok(int PP, int H, int K);
int tc_patch_fwd(const void* words, const void* w_bf16, const float* bias, const float* pos, const void* pos_bf16, void* out, int B, int PP, int Tn,
                 int off, int H, int K, cudaStream_t st);
size_t tc_patch_wgrad_ws_bytes(int B, int PP, int H, int K);
int tc_patch_wgrad(const void* dout, const void* words, float* dw, float* dbias, void* ws, size_t ws_bytes, int B, int Tn, int PP,
                   int has_cls, int H, int K, cudaStream_t st);

}  // namespace vitb
