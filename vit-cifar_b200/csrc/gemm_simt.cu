// gemm_simt.cu — fp32-FFMA GEMM with the shared epilogue.  Used for
//   * the fp32 check mode (bit-reproducible, sequential-k accumulation, no tensor cores),
//   * shapes the tcgen05 kernel does not take (head: N = num_classes; patch embedding: K = 48/192),
// and the patch-embedding front end (vit.py:66-70, 79-89) built on it.
#include "common.cuh"
#include "gemm_internal.h"

namespace vitb {

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }

template <typename TO> __device__ __forceinline__ void stf(TO* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// one output element through the shared epilogue (TO = activation type)
template <typename TO>
__device__ __forceinline__ void epi_apply(const EpiParams& e, int row, int col, float v, size_t raw_off) {
  if (e.mode == EPI_RAW_F32) {
    ((float*)e.out)[raw_off + (size_t)row * e.ldc + col] = v;
    return;
  }
  int prow = row;
  if (e.rm_group > 0) prow = (row / e.rm_group) * e.rm_stride + e.rm_offset + row % e.rm_group;
  const size_t off = (size_t)prow * e.ldc + col;
  if (e.mode == EPI_FWD) {
    if (e.bias) v += __ldg(e.bias + col);
    if (e.pos) v += __ldg(e.pos + (size_t)(prow % e.rm_stride) * e.ldc + col);
    if (e.preact) stf<TO>((TO*)e.preact + off, v);
    if (e.gelu) v = gelu_f(v);
    if (e.residual) v += ldf<TO>((const TO*)e.residual + off);
    if (e.out_f32) ((float*)e.out)[off] = v;
    else stf<TO>((TO*)e.out + off, v);
  } else {  // EPI_DGRAD
    if (e.aux) v *= gelu_grad_f(ldf<TO>((const TO*)e.aux + off));
    stf<TO>((TO*)e.out + off, v);
  }
}

// words(m, f) of an NCHW fp32 image: vit.py:79-89
__device__ __forceinline__ float gather_word(const float* __restrict__ img, int m, int f, int S, int P) {
  const int ps = S / P;
  const int c = f % 3, kw = (f / 3) % ps, kh = f / (3 * ps);
  const int pw = m % P, ph = (m / P) % P, b = m / (P * P);
  return __ldg(img + (((int64_t)b * 3 + c) * S + ph * ps + kh) * S + pw * ps + kw);
}

constexpr int SB = 64;   // block tile (M and N)
constexpr int SK = 16;   // k tile

template <typename TA, typename TB, typename TO>
__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtGemmArgs g) {
  pdl_trigger();
  pdl_wait();
  __shared__ float As[SK][SB + 4];
  __shared__ float Bs[SK][SB + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SB, n0 = blockIdx.x * SB;
  const int kper = (g.K + gridDim.z - 1) / gridDim.z;
  const int kbeg = blockIdx.z * kper;
  const int kend = min(g.K, kbeg + kper);
  const TA* __restrict__ A = (const TA*)g.a;
  const TB* __restrict__ B = (const TB*)g.b;
  const bool a_kfast = (g.a_sk == 1), b_kfast = (g.b_sk == 1);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += SK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int kk, mm;
      if (a_kfast) { kk = idx & (SK - 1); mm = idx >> 4; } else { mm = idx & (SB - 1); kk = idx >> 6; }
      const int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < g.M && gk < kend) {
        if (g.a_gather) {
          v = gather_word((const float*)g.a, gm, gk, g.gather_S, g.gather_P);
        } else {
        int64_t koff;
        if (g.a_kgroup > 0) koff = (int64_t)(gk / g.a_kgroup) * g.a_kgroup_stride + (int64_t)(gk % g.a_kgroup) * g.a_sk + g.a_koff;
        else koff = (int64_t)gk * g.a_sk;
        v = ldf<TA>(A + (int64_t)gm * g.a_sm + koff);
        }
      }
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int kk, nn;
      if (b_kfast) { kk = idx & (SK - 1); nn = idx >> 4; } else { nn = idx & (SB - 1); kk = idx >> 6; }
      const int gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < g.N && gk < kend) {
        if (g.b_gather) v = gather_word((const float*)g.b, gk, gn, g.gather_S, g.gather_P);
        else v = ldf<TB>(B + (int64_t)gk * g.b_sk + (int64_t)gn * g.b_sn);
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const size_t raw_off = (size_t)blockIdx.z * g.M * g.e.ldc;
  const int col0 = n0 + tx * 4;
  // fast path: plain activation output of 4 consecutive, aligned columns -> one 8/16-byte store per row
  const bool vec = (g.e.mode == EPI_FWD) && !g.e.preact && !g.e.gelu && !g.e.residual && !g.e.out_f32 && (col0 + 3 < g.N) && (g.e.ldc % 4 == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= g.M) continue;
    if (vec) {
      int prow = row;
      if (g.e.rm_group > 0) prow = (row / g.e.rm_group) * g.e.rm_stride + g.e.rm_offset + row % g.e.rm_group;
      float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      if (g.e.bias) { const float4 b = __ldg(reinterpret_cast<const float4*>(g.e.bias + col0)); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
      if (g.e.pos) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(g.e.pos + (size_t)(prow % g.e.rm_stride) * g.e.ldc + col0));
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      }
      st4((TO*)g.e.out + (size_t)prow * g.e.ldc + col0, v);
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + j;
      if (col < g.N) epi_apply<TO>(g.e, row, col, acc[i][j], raw_off);
    }
  }
}

// dt codes for operands: 0 = fp32, 1 = bf16
int simt_gemm_launch(const SimtGemmArgs& g, int a_dt, int b_dt, int o_dt, int splits, cudaStream_t st) {
  VITB_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "simt gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  dim3 grid(ceil_div(g.N, SB), ceil_div(g.M, SB), splits);
#define L(TA, TB, TO) VITB_LAUNCH((gemm_simt_kernel<TA, TB, TO>), grid, 256, 0, st, g)
  const int key = a_dt * 4 + b_dt * 2 + o_dt;
  switch (key) {
    case 0: L(float, float, float); break;
    case 1: L(float, float, bf16); break;
    case 2: L(float, bf16, float); break;
    case 3: L(float, bf16, bf16); break;
    case 4: L(bf16, float, float); break;
    case 5: L(bf16, float, bf16); break;
    case 6: L(bf16, bf16, float); break;
    default: L(bf16, bf16, bf16); break;
  }
#undef L
  VITB_LAUNCH_OK();
  return 0;
}

// small column sum (any cols): out[c] = sum_r x[r*ld + c]; one thread per column, fixed order
template <typename T>
__global__ void colsum_small_kernel(const T* __restrict__ x, float* __restrict__ out, int rows, int cols, int64_t ld) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += ldf<T>(x + (int64_t)r * ld + c);
  out[c] = s;
}

int colsum_small_launch(const void* x, float* out, int rows, int cols, int64_t ld, int x_dt, cudaStream_t st) {
  if (x_dt == VITB_BF16) VITB_LAUNCH((colsum_small_kernel<bf16>), ceil_div(cols, 128), 128, 0, st, (const bf16*)x, out, rows, cols, ld);
  else VITB_LAUNCH((colsum_small_kernel<float>), ceil_div(cols, 128), 128, 0, st, (const float*)x, out, rows, cols, ld);
  VITB_LAUNCH_OK();
  return 0;
}

// split count for a reduction of length K with `tiles` output tiles: fill ~2 waves, >= 64 k per split
int simt_pick_splits(int tiles, int K) {
  int s = (2 * kNumSMs) / (tiles > 0 ? tiles : 1);
  const int maxs = K / 64 > 0 ? K / 64 : 1;
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return s;
}

// ---------------------------------------------------------------------------------------------
// patch embedding front end
// ---------------------------------------------------------------------------------------------
// out[b,0,:] = cls + pos[0]   (vit.py:69-70)
template <typename T>
__global__ void cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, T* __restrict__ out,
                                int B, int Tn, int H) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < H; c += blockDim.x) Act<T>::st(out + (size_t)b * Tn * H + c, cls[c] + pos[c]);
}

// dcls = dpos[0]; dbias[c] = sum_{t >= has_cls} dpos[t][c]
__global__ void patch_bias_cls_kernel(const float* __restrict__ dpos, float* __restrict__ dbias, float* __restrict__ dcls,
                                      int Tn, int H, int has_cls) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  float s = 0.f;
  for (int t = has_cls; t < Tn; ++t) s += dpos[(size_t)t * H + c];
  dbias[c] = s;
  if (has_cls && dcls) dcls[c] = dpos[c];
}

// part[s][j] = sum over the s-th slice of the batch of dout[b][j], j over T*H (fixed order; finalize sums the slices)
template <typename T>
__global__ void batch_sum_kernel(const T* __restrict__ x, float* __restrict__ part, int B, int64_t n) {
  pdl_trigger();
  pdl_wait();
  const int64_t j4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j4 >= n) return;
  const int per = (B + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * per, b1 = min(B, b0 + per);
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  int b = b0;
  for (; b + 2 <= b1; b += 2) {
    const float4 v0 = ld4(x + (int64_t)b * n + j4);
    const float4 v1 = ld4(x + (int64_t)(b + 1) * n + j4);
    s0.x += v0.x; s0.y += v0.y; s0.z += v0.z; s0.w += v0.w;
    s1.x += v1.x; s1.y += v1.y; s1.z += v1.z; s1.w += v1.w;
  }
  if (b < b1) {
    const float4 v0 = ld4(x + (int64_t)b * n + j4);
    s0.x += v0.x; s0.y += v0.y; s0.z += v0.z; s0.w += v0.w;
  }
  *reinterpret_cast<float4*>(part + (int64_t)blockIdx.y * n + j4) = make_float4(s0.x + s1.x, s0.y + s1.y, s0.z + s1.z, s0.w + s1.w);
}

static int batch_sum_slices(int B) { return B >= 256 ? 16 : (B >= 32 ? 4 : 1); }

// words[m][f..f+7] (bf16) for the tensor-core path: 8 consecutive features per thread (vit.py:79-89).  The same launch also does
// the two small jobs of the stem that depend on nothing else: the bf16 copy of pos_emb the GEMM epilogue reads through TMA
// (pos_bf16 != nullptr) and the B cls rows out[b, 0, :] = cls + pos[0] (cls_out != nullptr; rows the GEMM does not write) —
// 4 launches -> 2 for the patch embedding.
__global__ void words_bf16_kernel(const float* __restrict__ img, bf16* __restrict__ words, int B, int S, int P, const float* __restrict__ pos,
                                  bf16* __restrict__ pos_bf16, int64_t n_pos, const float* __restrict__ cls, bf16* __restrict__ cls_out, int Tn,
                                  int H) {
  pdl_trigger();
  pdl_wait();
  const int ps = S / P;
  const int K = ps * ps * 3, K8 = K / 8;
  const int64_t total = (int64_t)B * P * P * K8;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < total; i += nthreads) {
    const int f0 = (int)(i % K8) * 8;
    const int m = (int)(i / K8);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = gather_word(img, m, f0 + j, S, P);
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(words + (int64_t)m * K + f0) = u;
  }
  if (pos_bf16 != nullptr) {
    for (int64_t i = tid; i < n_pos; i += nthreads) pos_bf16[i] = __float2bfloat16_rn(pos[i]);
  }
  if (cls_out != nullptr) {
    for (int64_t i = tid; i < (int64_t)B * H; i += nthreads) {
      const int b = (int)(i / H), c = (int)(i % H);
      cls_out[(size_t)b * Tn * H + c] = __float2bfloat16_rn(cls[c] + pos[c]);
    }
  }
}

}  // namespace vitb

using namespace vitb;

extern "C" {

static bool patch_tc_path(const void* words, const void* w_act, int P, int S, int H, int dt) {
  if (dt != VITB_BF16 || words == nullptr || w_act == nullptr || P <= 0 || S % P) return false;
  const int ps = S / P;
  return tc_patch_ok(P * P, H, ps * ps * 3);
}

size_t vitb_patch_embed_fwd_ws_bytes(int B, int S, int P, int H, int dt) {
  (void)B; (void)S;
  if (dt != VITB_BF16 || P <= 0) return 0;
  return align_up((size_t)(P * P + 1) * H * sizeof(bf16), 256);  // bf16 copy of pos_emb for the GEMM epilogue's TMA loads
}

int vitb_patch_embed_fwd(const float* img, const float* w, const void* w_act, const float* bias, const float* cls, const float* pos,
                         void* out, void* words, void* ws, size_t ws_bytes, int B, int S, int P, int H, int has_cls, int dt, void* stream) {
  VITB_REQUIRE(img && w && bias && pos && out, "patch_embed_fwd: null pointer");
  VITB_REQUIRE(B > 0 && P > 0 && S % P == 0 && H % 4 == 0, "patch_embed_fwd: bad shape S=%d P=%d H=%d", S, P, H);
  VITB_REQUIRE(!has_cls || cls, "patch_embed_fwd: has_cls without cls pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int ps = S / P, K = ps * ps * 3, PP = P * P, Tn = PP + (has_cls ? 1 : 0);
  if (patch_tc_path(words, w_act, P, S, H, dt)) {
    // tensor-core path: patch gather (bf16 words, kept for the backward wgrad) -> tcgen05 GEMM whose epilogue adds bias and
    // pos_emb[token] and stores into the token rows of (B, T, H) -> the B cls rows
    int blocks = (int)ceil_div64((int64_t)B * PP * (K / 8), 256);
    if (blocks > 16 * kNumSMs) blocks = 16 * kNumSMs;
    const void* pos_bf16 = nullptr;
    if (PP % 32 == 0 && ws != nullptr && ws_bytes >= vitb_patch_embed_fwd_ws_bytes(B, S, P, H, dt)) pos_bf16 = ws;
    VITB_LAUNCH((words_bf16_kernel), blocks, 256, 0, st, img, (bf16*)words, B, S, P, pos, (bf16*)const_cast<void*>(pos_bf16), (int64_t)Tn * H,
                has_cls ? cls : nullptr, has_cls ? (bf16*)out : nullptr, Tn, H);
    VITB_LAUNCH_OK();
    return tc_patch_fwd(words, w_act, bias, pos, pos_bf16, out, B, PP, Tn, has_cls ? 1 : 0, H, K, st);
  }
  SimtGemmArgs g = {};
  g.a = img; g.b = w;
  g.M = B * PP; g.N = H; g.K = K;
  g.a_gather = 1; g.gather_S = S; g.gather_P = P;  // A(m,k) = words(m,k) read straight from the image
  g.b_sk = 1; g.b_sn = K;                          // B(k,n) = w[n][k]
  g.e.mode = EPI_FWD; g.e.bias = bias; g.e.out = out; g.e.ldc = H;
  g.e.rm_group = PP; g.e.rm_stride = Tn; g.e.rm_offset = has_cls ? 1 : 0; g.e.pos = pos;
  int rc = simt_gemm_launch(g, VITB_F32, VITB_F32, dt, 1, st);
  if (rc) return rc;
  if (has_cls) {
    if (dt == VITB_BF16) VITB_LAUNCH((cls_rows_kernel<bf16>), B, 128, 0, st, cls, pos, (bf16*)out, B, Tn, H);
    else VITB_LAUNCH((cls_rows_kernel<float>), B, 128, 0, st, cls, pos, (float*)out, B, Tn, H);
    VITB_LAUNCH_OK();
  }
  return 0;
}

static size_t patch_bwd_simt_part_bytes(int B, int PP, int H, int K) {
  const int tiles = ceil_div(H, SB) * ceil_div(K, SB);
  return align_up((size_t)simt_pick_splits(tiles, B * PP) * H * K * sizeof(float), 256);
}

size_t vitb_patch_embed_bwd_ws_bytes(int B, int S, int P, int H, int has_cls) {
  if (P <= 0 || S % P) return 0;
  const int ps = S / P, K = ps * ps * 3, PP = P * P;
  const int Tn = PP + (has_cls ? 1 : 0);
  size_t wg = patch_bwd_simt_part_bytes(B, PP, H, K);
  if (tc_patch_ok(PP, H, K)) {
    const size_t t = tc_patch_wgrad_ws_bytes(B, PP, H, K);
    if (t > wg) wg = t;
  }
  return wg + align_up((size_t)batch_sum_slices(B) * Tn * H * sizeof(float), 256);
}

int vitb_patch_embed_bwd(const float* img, const void* words, const void* dout, float* dw, float* dbias, float* dcls, float* dpos,
                         void* ws, size_t ws_bytes, int B, int S, int P, int H, int has_cls, int dt, void* stream) {
  VITB_REQUIRE(img && dout && dw && dbias && dpos && ws, "patch_embed_bwd: null pointer");
  VITB_REQUIRE(B > 0 && P > 0 && S % P == 0 && H % 4 == 0, "patch_embed_bwd: bad shape");
  VITB_REQUIRE(ws_bytes >= vitb_patch_embed_bwd_ws_bytes(B, S, P, H, has_cls), "patch_embed_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int ps = S / P, K = ps * ps * 3, PP = P * P, Tn = PP + (has_cls ? 1 : 0);
  const bool tc = dt == VITB_BF16 && words != nullptr && tc_patch_ok(PP, H, K);
  const size_t wg_bytes = vitb_patch_embed_bwd_ws_bytes(B, S, P, H, has_cls) - align_up((size_t)batch_sum_slices(B) * Tn * H * sizeof(float), 256);
  // 1) dpos = sum_b dout[b]; dcls (and, on the SIMT path, dbias) from it
  {
    const int64_t n = (int64_t)Tn * H;
    const int slices = batch_sum_slices(B);
    float* bpart = slices > 1 ? (float*)((char*)ws + wg_bytes) : dpos;
    const dim3 grid((unsigned)ceil_div64(n / 4, 128), slices);
    if (dt == VITB_BF16) VITB_LAUNCH((batch_sum_kernel<bf16>), grid, 128, 0, st, (const bf16*)dout, bpart, B, n);
    else VITB_LAUNCH((batch_sum_kernel<float>), grid, 128, 0, st, (const float*)dout, bpart, B, n);
    VITB_LAUNCH_OK();
    if (slices > 1) {
      (void)::vitb::launch_finalize(bpart, slices, n, dpos, nullptr, nullptr, 1, st);
      VITB_LAUNCH_OK();
    }
    VITB_LAUNCH((patch_bias_cls_kernel), ceil_div(H, 128), 128, 0, st, dpos, dbias, dcls, Tn, H, has_cls ? 1 : 0);
    VITB_LAUNCH_OK();
  }
  // 2) dW[h][k] = sum_m dout[phys(m)][h] * words(m,k)
  if (tc) return tc_patch_wgrad(dout, words, dw, dbias, ws, wg_bytes, B, Tn, PP, has_cls, H, K, st);  // (rewrites dbias, same value)
  float* part = (float*)ws;
  const int tiles = ceil_div(H, SB) * ceil_div(K, SB);
  const int splits = simt_pick_splits(tiles, B * PP);
  SimtGemmArgs g = {};
  g.a = dout; g.b = img;
  g.M = H; g.N = K; g.K = B * PP;
  g.a_sm = 1; g.a_sk = H; g.a_kgroup = PP; g.a_kgroup_stride = (int64_t)Tn * H; g.a_koff = (int64_t)(has_cls ? 1 : 0) * H;
  g.b_gather = 1; g.gather_S = S; g.gather_P = P;
  g.e.mode = EPI_RAW_F32; g.e.ldc = K; g.e.out = splits > 1 ? part : dw;
  int rc = simt_gemm_launch(g, dt, VITB_F32, VITB_F32, splits, st);
  if (rc) return rc;
  if (splits > 1) {
    const int64_t n = (int64_t)H * K;
    (void)::vitb::launch_finalize(part, splits, n, dw, nullptr, nullptr, 1, st);
    VITB_LAUNCH_OK();
  }
  return 0;
}

}  // extern "C"
