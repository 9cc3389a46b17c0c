// attention.cu — fused short-sequence attention (layers.py:92-101), forward and backward.
//
// T is 17 or 65 (<= 128), head_dim 32 (64 for the scaled config): the whole (T x T) score tile of one
// (image, head) lives on chip.  bf16 path: one CTA per (image, head), one warp per 16 query rows,
// QKᵀ and PV on mma.sync.m16n8k16 bf16 tensor-core tiles fed by ldmatrix from padded shared memory
// (tcgen05's 128-row tiles do not fit a 65x65x32 problem), softmax in registers with quad shuffles;
// only the per-row log-sum-exp is kept for backward, which recomputes P.  fp32 path (check mode):
// plain FFMA with the score tile in shared memory.
#include "common.cuh"

namespace vitb {

// ---------------------------------------------------------------------------------------------
// mma / ldmatrix wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

constexpr float kLog2e = 1.4426950408889634f;

// load a (T x D) head slice of the packed (B,T,3H) tensor into padded smem, zero rows >= T
template <int D, int TP>
__device__ __forceinline__ void load_head_tile(bf16* dst, const bf16* src, int64_t row_stride, int T, int tid, int nthr) {
  constexpr int LD = D + 8, CH = D / 8;
  for (int idx = tid; idx < TP * CH; idx += nthr) {
    const int r = idx / CH, c = idx % CH;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < T) v = *reinterpret_cast<const uint4*>(src + (int64_t)r * row_stride + c * 8);
    *reinterpret_cast<uint4*>(dst + r * LD + c * 8) = v;
  }
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// asynchronous version: rows < T only (the padding rows are zeroed once per kernel and never written again)
template <int D, int TP>
__device__ __forceinline__ void load_head_tile_async(bf16* dst, const bf16* src, int64_t row_stride, int T, int tid, int nthr) {
  constexpr int LD = D + 8, CH = D / 8;
  for (int idx = tid; idx < T * CH; idx += nthr) {
    const int r = idx / CH, c = idx % CH;
    cp_async16(dst + r * LD + c * 8, src + (int64_t)r * row_stride + c * 8);
  }
}
template <int D, int TP>
__device__ __forceinline__ void zero_pad_rows(bf16* dst, int T, int tid, int nthr) {
  constexpr int LD = D + 8;
  for (int idx = tid; idx < (TP - T) * LD / 2; idx += nthr) reinterpret_cast<uint32_t*>(dst + T * LD)[idx] = 0u;
}

// S[16 x TP] = Q_rows(16w..) · Kᵀ   (raw, unscaled) -> s[j][4], j = key tile of 8
template <int D, int NT16>
__device__ __forceinline__ void qk_tile(float (&s)[2 * NT16][4], const bf16* sA, const bf16* sB, int warp, int lane) {
  constexpr int LD = D + 8;
  uint32_t a[D / 16][4];
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) ldsm_x4(a[kk], sA + (16 * warp + (lane & 15)) * LD + kk * 16 + (lane >> 4) * 8);
#pragma unroll
  for (int j = 0; j < 2 * NT16; ++j) {
    s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int k2 = 0; k2 < D / 32; ++k2) {
      uint32_t b[4];
      ldsm_x4(b, sB + (8 * j + (lane & 7)) * LD + k2 * 32 + (lane >> 3) * 8);
      mma_bf16(s[j], a[2 * k2], b[0], b[1]);
      mma_bf16(s[j], a[2 * k2 + 1], b[2], b[3]);
    }
  }
}

// acc[16 x D] += P(regs, 16 x TP) · B(TP x D) with B rows = reduction index (ldmatrix.trans)
template <int D, int NT16>
__device__ __forceinline__ void pv_tile(float (&acc)[D / 8][4], const float (&p)[2 * NT16][4], const bf16* sB, int lane) {
  constexpr int LD = D + 8;
#pragma unroll
  for (int kk = 0; kk < NT16; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int jp = 0; jp < D / 16; ++jp) {
      uint32_t b[4];
      ldsm_x4_t(b, sB + (16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + jp * 16 + (lane >> 4) * 8);
      mma_bf16(acc[2 * jp], a, b[0], b[1]);
      mma_bf16(acc[2 * jp + 1], a, b[2], b[3]);
    }
  }
}

// write a warp's 16 x D fp32 fragment tile as bf16 into its own 16 smem rows, then stream it out coalesced
template <int D>
__device__ __forceinline__ void store_rows16(const float (&acc)[D / 8][4], bf16* sTile /* row 0 of this warp */, bf16* gdst,
                                             int64_t g_row_stride, int rows_valid, int lane) {
  constexpr int LD = D + 8, CH = D / 8;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int jn = 0; jn < D / 8; ++jn) {
    *reinterpret_cast<uint32_t*>(sTile + g * LD + jn * 8 + 2 * t) = pack_bf16x2(acc[jn][0], acc[jn][1]);
    *reinterpret_cast<uint32_t*>(sTile + (g + 8) * LD + jn * 8 + 2 * t) = pack_bf16x2(acc[jn][2], acc[jn][3]);
  }
  __syncwarp();
  for (int idx = lane; idx < 16 * CH; idx += 32) {
    const int r = idx / CH, c = idx % CH;
    if (r < rows_valid) *reinterpret_cast<uint4*>(gdst + (int64_t)r * g_row_stride + c * 8) = *reinterpret_cast<const uint4*>(sTile + r * LD + c * 8);
  }
}

// ---------------------------------------------------------------------------------------------
// bf16 forward: persistent CTAs walk (image, head) items; the next item's Q/K/V tiles stream in with cp.async
// (double buffer) while the current one is computed
// ---------------------------------------------------------------------------------------------
template <int D, int NT16>
__global__ void __launch_bounds__(32 * NT16, (NT16 <= 2 ? 8 : (NT16 <= 5 ? 4 : 1)))
    attn_fwd_bf16_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o, float* __restrict__ lse, float* __restrict__ attn_map,
                         int n_items, int T, int heads, float scale) {
  constexpr int TP = 16 * NT16, LD = D + 8, TILE = TP * LD;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  bf16* sbuf = reinterpret_cast<bf16*>(smem_attn);  // [2][3][TILE]
  const int Hd = heads * D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x;
  for (int i = 0; i < 6; ++i) zero_pad_rows<D, TP>(sbuf + i * TILE, T, tid, nthr);

  auto issue = [&](int item, int buf) {
    const int b = item / heads, h = item % heads;
    const bf16* base = qkv + (int64_t)b * T * 3 * Hd + h * D;
    bf16* dst = sbuf + buf * 3 * TILE;
    load_head_tile_async<D, TP>(dst, base, 3 * Hd, T, tid, nthr);
    load_head_tile_async<D, TP>(dst + TILE, base + Hd, 3 * Hd, T, tid, nthr);
    load_head_tile_async<D, TP>(dst + 2 * TILE, base + 2 * Hd, 3 * Hd, T, tid, nthr);
  };

  int item = blockIdx.x;
  if (item < n_items) issue(item, 0);
  cp_async_commit();
  const int g = lane >> 2, t = lane & 3;
  const float sl2 = scale * kLog2e;
  for (int it = 0; item < n_items; item += gridDim.x, ++it) {
    const int cur = it & 1;
    const int nxt = item + gridDim.x;
    if (nxt < n_items) issue(nxt, cur ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    bf16* sQ = sbuf + cur * 3 * TILE;
    bf16* sK = sQ + TILE;
    bf16* sV = sK + TILE;
    const int b = item / heads, h = item % heads;

    float s[2 * NT16][4];
    qk_tile<D, NT16>(s, sQ, sK, warp, lane);
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2 * NT16; ++j) {
      const int c = 8 * j + 2 * t;
      if (c >= T) s[j][0] = s[j][2] = -INFINITY;
      if (c + 1 >= T) s[j][1] = s[j][3] = -INFINITY;
      m0 = fmaxf(m0, fmaxf(s[j][0], s[j][1]));
      m1 = fmaxf(m1, fmaxf(s[j][2], s[j][3]));
    }
    m0 = quad_max(m0);
    m1 = quad_max(m1);
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int j = 0; j < 2 * NT16; ++j) {
      s[j][0] = exp2f((s[j][0] - m0) * sl2);
      s[j][1] = exp2f((s[j][1] - m0) * sl2);
      s[j][2] = exp2f((s[j][2] - m1) * sl2);
      s[j][3] = exp2f((s[j][3] - m1) * sl2);
      sum0 += s[j][0] + s[j][1];
      sum1 += s[j][2] + s[j][3];
    }
    sum0 = quad_sum(sum0);
    sum1 = quad_sum(sum1);
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    const int r0 = 16 * warp + g, r1 = r0 + 8;
    if (t == 0) {
      float* l = lse + ((int64_t)b * heads + h) * T;
      if (r0 < T) l[r0] = m0 * scale + logf(sum0);
      if (r1 < T) l[r1] = m1 * scale + logf(sum1);
    }
#pragma unroll
    for (int j = 0; j < 2 * NT16; ++j) {
      s[j][0] *= inv0; s[j][1] *= inv0; s[j][2] *= inv1; s[j][3] *= inv1;
    }
    if (attn_map != nullptr) {  // save_attn_map protocol (layers.py:99-100)
      float* am = attn_map + ((int64_t)b * heads + h) * T * T;
#pragma unroll
      for (int j = 0; j < 2 * NT16; ++j) {
        const int c = 8 * j + 2 * t;
        if (r0 < T) { if (c < T) am[(int64_t)r0 * T + c] = s[j][0]; if (c + 1 < T) am[(int64_t)r0 * T + c + 1] = s[j][1]; }
        if (r1 < T) { if (c < T) am[(int64_t)r1 * T + c] = s[j][2]; if (c + 1 < T) am[(int64_t)r1 * T + c + 1] = s[j][3]; }
      }
    }
    float acc[D / 8][4];
#pragma unroll
    for (int jn = 0; jn < D / 8; ++jn) acc[jn][0] = acc[jn][1] = acc[jn][2] = acc[jn][3] = 0.f;
    pv_tile<D, NT16>(acc, s, sV, lane);
    // this warp's Q rows are dead (fragments already in registers, nobody else reads them): reuse as staging.
    // (padding rows of Q may now hold garbage: they only feed score rows >= T, which are never stored)
    const int rows_valid = min(16, T - 16 * warp);
    if (rows_valid > 0)
      store_rows16<D>(acc, sQ + 16 * warp * LD, o + ((int64_t)b * T + 16 * warp) * Hd + h * D, Hd, rows_valid, lane);
    __syncthreads();  // everyone is done with this buffer before the next iteration refills it
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// bf16 backward
// ---------------------------------------------------------------------------------------------
// acc[16 keys x D] += Aᵀ-tile from sPS (stored [query][key]) · sB (stored [query][D]); reduction over queries
template <int D, int NT16>
__device__ __forceinline__ void tn_tile(float (&acc)[D / 8][4], const bf16* sPS, const bf16* sB, int warp, int lane) {
  constexpr int LD = D + 8, LP = 16 * NT16 + 8;
#pragma unroll
  for (int kq = 0; kq < NT16; ++kq) {
    uint32_t a[4];
    ldsm_x4_t(a, sPS + (16 * kq + (lane & 7) + ((lane >> 4) & 1) * 8) * LP + 16 * warp + ((lane >> 3) & 1) * 8);
#pragma unroll
    for (int jp = 0; jp < D / 16; ++jp) {
      uint32_t b[4];
      ldsm_x4_t(b, sB + (16 * kq + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + jp * 16 + (lane >> 4) * 8);
      mma_bf16(acc[2 * jp], a, b[0], b[1]);
      mma_bf16(acc[2 * jp + 1], a, b[2], b[3]);
    }
  }
}

template <int D, int NT16>
__global__ void __launch_bounds__(32 * NT16, (NT16 <= 2 ? 4 : (NT16 <= 5 ? 2 : 1)))
    attn_bwd_bf16_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o, const float* __restrict__ lse,
                         bf16* __restrict__ dqkv, int n_items, int T, int heads, float scale) {
  constexpr int TP = 16 * NT16, LD = D + 8, LP = TP + 8, TILE = TP * LD;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  bf16* sbuf = reinterpret_cast<bf16*>(smem_attn);  // [2][4][TILE]  (Q, K, V, dO)
  bf16* sP = sbuf + 8 * TILE;                        // [TP][LP]
  bf16* sdS = sP + TP * LP;                          // [TP][LP]
  float* sLse = reinterpret_cast<float*>(sdS + TP * LP);  // [2][TP]
  const int Hd = heads * D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x;
  for (int i = 0; i < 8; ++i) zero_pad_rows<D, TP>(sbuf + i * TILE, T, tid, nthr);

  auto issue = [&](int item, int buf) {
    const int b = item / heads, h = item % heads;
    const bf16* base = qkv + (int64_t)b * T * 3 * Hd + h * D;
    bf16* dst = sbuf + buf * 4 * TILE;
    load_head_tile_async<D, TP>(dst, base, 3 * Hd, T, tid, nthr);
    load_head_tile_async<D, TP>(dst + TILE, base + Hd, 3 * Hd, T, tid, nthr);
    load_head_tile_async<D, TP>(dst + 2 * TILE, base + 2 * Hd, 3 * Hd, T, tid, nthr);
    load_head_tile_async<D, TP>(dst + 3 * TILE, d_o + (int64_t)b * T * Hd + h * D, Hd, T, tid, nthr);
    for (int r = tid; r < T; r += nthr) cp_async4(sLse + buf * TP + r, lse + ((int64_t)b * heads + h) * T + r);
  };

  int item = blockIdx.x;
  if (item < n_items) issue(item, 0);
  cp_async_commit();
  const int g = lane >> 2, t = lane & 3;
  const int r0 = 16 * warp + g, r1 = r0 + 8;
  const float sl2 = scale * kLog2e;
  for (int it = 0; item < n_items; item += gridDim.x, ++it) {
    const int cur = it & 1;
    const int nxt = item + gridDim.x;
    if (nxt < n_items) issue(nxt, cur ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    bf16* sQ = sbuf + cur * 4 * TILE;
    bf16* sK = sQ + TILE;
    bf16* sV = sK + TILE;
    bf16* sdO = sV + TILE;
    const int b = item / heads, h = item % heads;

    float s[2 * NT16][4], dp[2 * NT16][4];
    qk_tile<D, NT16>(s, sQ, sK, warp, lane);    // S = Q Kᵀ
    qk_tile<D, NT16>(dp, sdO, sV, warp, lane);  // dP = dO Vᵀ
    const float l0 = r0 < T ? sLse[cur * TP + r0] * kLog2e : 0.f;
    const float l1 = r1 < T ? sLse[cur * TP + r1] * kLog2e : 0.f;
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int j = 0; j < 2 * NT16; ++j) {
      const int c = 8 * j + 2 * t;
      const bool v0 = c < T, v1 = c + 1 < T;
      s[j][0] = (v0 && r0 < T) ? exp2f(s[j][0] * sl2 - l0) : 0.f;
      s[j][1] = (v1 && r0 < T) ? exp2f(s[j][1] * sl2 - l0) : 0.f;
      s[j][2] = (v0 && r1 < T) ? exp2f(s[j][2] * sl2 - l1) : 0.f;
      s[j][3] = (v1 && r1 < T) ? exp2f(s[j][3] * sl2 - l1) : 0.f;
      d0 += s[j][0] * dp[j][0] + s[j][1] * dp[j][1];
      d1 += s[j][2] * dp[j][2] + s[j][3] * dp[j][3];
    }
    d0 = quad_sum(d0);
    d1 = quad_sum(d1);
#pragma unroll
    for (int j = 0; j < 2 * NT16; ++j) {
      const int c = 8 * j + 2 * t;
      // dS = P ∘ (dP − rowsum(P ∘ dP)) / sqrt(features)
      dp[j][0] = s[j][0] * (dp[j][0] - d0) * scale;
      dp[j][1] = s[j][1] * (dp[j][1] - d0) * scale;
      dp[j][2] = s[j][2] * (dp[j][2] - d1) * scale;
      dp[j][3] = s[j][3] * (dp[j][3] - d1) * scale;
      *reinterpret_cast<uint32_t*>(sP + r0 * LP + c) = pack_bf16x2(s[j][0], s[j][1]);
      *reinterpret_cast<uint32_t*>(sP + r1 * LP + c) = pack_bf16x2(s[j][2], s[j][3]);
      *reinterpret_cast<uint32_t*>(sdS + r0 * LP + c) = pack_bf16x2(dp[j][0], dp[j][1]);
      *reinterpret_cast<uint32_t*>(sdS + r1 * LP + c) = pack_bf16x2(dp[j][2], dp[j][3]);
    }
    // dQ = dS · K   (reduction over keys; K rows are the reduction index -> ldmatrix.trans)
    float dq[D / 8][4];
#pragma unroll
    for (int jn = 0; jn < D / 8; ++jn) dq[jn][0] = dq[jn][1] = dq[jn][2] = dq[jn][3] = 0.f;
    pv_tile<D, NT16>(dq, dp, sK, lane);
    __syncthreads();  // sP / sdS complete

    // this warp now owns key rows 16w..16w+15:  dV = Pᵀ · dO,  dK = dSᵀ · Q
    float dv[D / 8][4], dk[D / 8][4];
#pragma unroll
    for (int jn = 0; jn < D / 8; ++jn) {
      dv[jn][0] = dv[jn][1] = dv[jn][2] = dv[jn][3] = 0.f;
      dk[jn][0] = dk[jn][1] = dk[jn][2] = dk[jn][3] = 0.f;
    }
    tn_tile<D, NT16>(dv, sP, sdO, warp, lane);
    tn_tile<D, NT16>(dk, sdS, sQ, warp, lane);
    __syncthreads();  // everyone is done reading sQ/sK/sV/sdO (and sP/sdS): reuse own rows as staging

    const int rows_valid = min(16, T - 16 * warp);
    if (rows_valid > 0) {
      bf16* gd = dqkv + ((int64_t)b * T + 16 * warp) * 3 * Hd + h * D;
      store_rows16<D>(dq, sQ + 16 * warp * LD, gd, 3 * Hd, rows_valid, lane);
      store_rows16<D>(dk, sK + 16 * warp * LD, gd + Hd, 3 * Hd, rows_valid, lane);
      store_rows16<D>(dv, sV + 16 * warp * LD, gd + 2 * Hd, 3 * Hd, rows_valid, lane);
    }
    // staging dirtied the padding rows of Q, K and V in this buffer: restore the zeros the next item relies on
    __syncthreads();
    zero_pad_rows<D, TP>(sQ, T, tid, nthr);
    zero_pad_rows<D, TP>(sK, T, tid, nthr);
    zero_pad_rows<D, TP>(sV, T, tid, nthr);
    // (no barrier needed here: the next write into this buffer is the cp.async of rows < T two iterations later,
    //  and every read of the padding rows is behind that iteration's __syncthreads)
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// fp32 check-mode kernels (FFMA, score tile in shared memory)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    attn_fwd_f32_kernel(const float* __restrict__ qkv, float* __restrict__ o, float* __restrict__ lse, float* __restrict__ attn_map,
                        int T, int heads, int D, float scale) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  float* sQ = reinterpret_cast<float*>(smem_attn);
  float* sK = sQ + T * D;
  float* sV = sK + T * D;
  float* sS = sV + T * D;  // T x T
  const int b = blockIdx.x / heads, h = blockIdx.x % heads, Hd = heads * D;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const float* base = qkv + (int64_t)b * T * 3 * Hd + h * D;
  for (int i = tid; i < T * D; i += nthr) {
    const int r = i / D, c = i % D;
    sQ[i] = base[(int64_t)r * 3 * Hd + c];
    sK[i] = base[(int64_t)r * 3 * Hd + Hd + c];
    sV[i] = base[(int64_t)r * 3 * Hd + 2 * Hd + c];
  }
  __syncthreads();
  for (int i = tid; i < T * T; i += nthr) {
    const int r = i / T, c = i % T;
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc = fmaf(sQ[r * D + d], sK[c * D + d], acc);
    sS[i] = acc * scale;
  }
  __syncthreads();
  for (int r = tid; r < T; r += nthr) {
    float m = -INFINITY;
    for (int c = 0; c < T; ++c) m = fmaxf(m, sS[r * T + c]);
    float sum = 0.f;
    for (int c = 0; c < T; ++c) {
      const float e = expf(sS[r * T + c] - m);
      sS[r * T + c] = e;
      sum += e;
    }
    const float inv = 1.0f / sum;
    for (int c = 0; c < T; ++c) sS[r * T + c] *= inv;
    lse[((int64_t)b * heads + h) * T + r] = m + logf(sum);
  }
  __syncthreads();
  if (attn_map != nullptr)
    for (int i = tid; i < T * T; i += nthr) attn_map[((int64_t)b * heads + h) * T * T + i] = sS[i];
  for (int i = tid; i < T * D; i += nthr) {
    const int r = i / D, d = i % D;
    float acc = 0.f;
    for (int c = 0; c < T; ++c) acc = fmaf(sS[r * T + c], sV[c * D + d], acc);
    o[((int64_t)b * T + r) * Hd + h * D + d] = acc;
  }
}

__global__ void __launch_bounds__(128)
    attn_bwd_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ d_o, const float* __restrict__ lse,
                        float* __restrict__ dqkv, int T, int heads, int D, float scale) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  float* sQ = reinterpret_cast<float*>(smem_attn);
  float* sK = sQ + T * D;
  float* sV = sK + T * D;
  float* sdO = sV + T * D;
  float* sP = sdO + T * D;  // T x T
  float* sdS = sP + T * T;  // T x T
  const int b = blockIdx.x / heads, h = blockIdx.x % heads, Hd = heads * D;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const float* base = qkv + (int64_t)b * T * 3 * Hd + h * D;
  for (int i = tid; i < T * D; i += nthr) {
    const int r = i / D, c = i % D;
    sQ[i] = base[(int64_t)r * 3 * Hd + c];
    sK[i] = base[(int64_t)r * 3 * Hd + Hd + c];
    sV[i] = base[(int64_t)r * 3 * Hd + 2 * Hd + c];
    sdO[i] = d_o[((int64_t)b * T + r) * Hd + h * D + c];
  }
  __syncthreads();
  const float* l = lse + ((int64_t)b * heads + h) * T;
  for (int i = tid; i < T * T; i += nthr) {
    const int r = i / T, c = i % T;
    float acc = 0.f, dp = 0.f;
    for (int d = 0; d < D; ++d) {
      acc = fmaf(sQ[r * D + d], sK[c * D + d], acc);
      dp = fmaf(sdO[r * D + d], sV[c * D + d], dp);
    }
    sP[i] = expf(acc * scale - l[r]);
    sdS[i] = dp;
  }
  __syncthreads();
  for (int r = tid; r < T; r += nthr) {
    float dsum = 0.f;
    for (int c = 0; c < T; ++c) dsum = fmaf(sP[r * T + c], sdS[r * T + c], dsum);
    for (int c = 0; c < T; ++c) sdS[r * T + c] = sP[r * T + c] * (sdS[r * T + c] - dsum) * scale;
  }
  __syncthreads();
  float* gd = dqkv + (int64_t)b * T * 3 * Hd + h * D;
  for (int i = tid; i < T * D; i += nthr) {
    const int r = i / D, d = i % D;
    float q = 0.f, k = 0.f, v = 0.f;
    for (int c = 0; c < T; ++c) {
      q = fmaf(sdS[r * T + c], sK[c * D + d], q);    // dQ[r] = sum_c dS[r][c] K[c]
      k = fmaf(sdS[c * T + r], sQ[c * D + d], k);    // dK[r] = sum_c dS[c][r] Q[c]
      v = fmaf(sP[c * T + r], sdO[c * D + d], v);    // dV[r] = sum_c P[c][r] dO[c]
    }
    gd[(int64_t)r * 3 * Hd + d] = q;
    gd[(int64_t)r * 3 * Hd + Hd + d] = k;
    gd[(int64_t)r * 3 * Hd + 2 * Hd + d] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
template <int D, int NT16>
static int launch_fwd_bf16(const void* qkv, void* o, float* lse, float* am, int B, int T, int heads, float scale, cudaStream_t st) {
  constexpr int TP = 16 * NT16, LD = D + 8;
  constexpr size_t smem = (size_t)6 * TP * LD * sizeof(bf16);
  constexpr int per_sm = NT16 <= 2 ? 8 : (NT16 <= 5 ? 4 : 1);
  auto kern = attn_fwd_bf16_kernel<D, NT16>;
  static bool configured = false;
  if (!configured) {
    VITB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int items = B * heads;
  const int grid = items < kNumSMs * per_sm ? items : kNumSMs * per_sm;
  kern<<<grid, 32 * NT16, smem, st>>>((const bf16*)qkv, (bf16*)o, lse, am, items, T, heads, scale);
  VITB_LAUNCH_OK();
  return 0;
}
template <int D, int NT16>
static int launch_bwd_bf16(const void* qkv, const void* d_o, const float* lse, void* dqkv, int B, int T, int heads, float scale, cudaStream_t st) {
  constexpr int TP = 16 * NT16, LD = D + 8, LP = TP + 8;
  constexpr size_t smem = ((size_t)8 * TP * LD + (size_t)2 * TP * LP) * sizeof(bf16) + (size_t)2 * TP * sizeof(float);
  constexpr int per_sm = NT16 <= 2 ? 4 : (NT16 <= 5 ? 2 : 1);
  auto kern = attn_bwd_bf16_kernel<D, NT16>;
  static bool configured = false;
  if (!configured) {
    VITB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  const int items = B * heads;
  const int grid = items < kNumSMs * per_sm ? items : kNumSMs * per_sm;
  kern<<<grid, 32 * NT16, smem, st>>>((const bf16*)qkv, (const bf16*)d_o, lse, (bf16*)dqkv, items, T, heads, scale);
  VITB_LAUNCH_OK();
  return 0;
}

#define VITB_ATTN_DISPATCH(FN, ...)                                                   \
  do {                                                                                \
    const int nt = (T + 15) / 16;                                                     \
    if (d == 32) {                                                                    \
      if (nt <= 1) return FN<32, 1>(__VA_ARGS__);                                     \
      if (nt <= 2) return FN<32, 2>(__VA_ARGS__);                                     \
      if (nt <= 4) return FN<32, 4>(__VA_ARGS__);                                     \
      if (nt <= 5) return FN<32, 5>(__VA_ARGS__);                                     \
      return FN<32, 8>(__VA_ARGS__);                                                  \
    } else {                                                                          \
      if (nt <= 1) return FN<64, 1>(__VA_ARGS__);                                     \
      if (nt <= 2) return FN<64, 2>(__VA_ARGS__);                                     \
      if (nt <= 4) return FN<64, 4>(__VA_ARGS__);                                     \
      if (nt <= 5) return FN<64, 5>(__VA_ARGS__);                                     \
      return FN<64, 8>(__VA_ARGS__);                                                  \
    }                                                                                 \
  } while (0)

static int set_dyn_smem(const void* fn, size_t smem) {
  if (smem > 48 * 1024) VITB_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return 0;
}

}  // namespace vitb

using namespace vitb;

extern "C" {

int vitb_attn_fwd(const void* qkv, void* o, float* lse, float* attn_map, int B, int T, int heads, int d, float scale, int dt, void* stream) {
  VITB_REQUIRE(qkv && o && lse, "attn_fwd: null pointer");
  VITB_REQUIRE(B > 0 && T > 0 && T <= 128 && heads > 0 && (d == 32 || d == 64), "attn_fwd: unsupported shape B=%d T=%d heads=%d d=%d (T<=128, d in {32,64})", B, T, heads, d);
  cudaStream_t st = (cudaStream_t)stream;
  if (dt == VITB_BF16) {
    VITB_ATTN_DISPATCH(launch_fwd_bf16, qkv, o, lse, attn_map, B, T, heads, scale, st);
  }
  const size_t smem = ((size_t)3 * T * d + (size_t)T * T) * sizeof(float);
  if (set_dyn_smem((const void*)attn_fwd_f32_kernel, smem)) return -1;
  attn_fwd_f32_kernel<<<B * heads, 128, smem, st>>>((const float*)qkv, (float*)o, lse, attn_map, T, heads, d, scale);
  VITB_LAUNCH_OK();
  return 0;
}

int vitb_attn_bwd(const void* qkv, const void* d_o, const float* lse, void* dqkv, int B, int T, int heads, int d, float scale, int dt, void* stream) {
  VITB_REQUIRE(qkv && d_o && lse && dqkv, "attn_bwd: null pointer");
  VITB_REQUIRE(B > 0 && T > 0 && T <= 128 && heads > 0 && (d == 32 || d == 64), "attn_bwd: unsupported shape B=%d T=%d heads=%d d=%d (T<=128, d in {32,64})", B, T, heads, d);
  cudaStream_t st = (cudaStream_t)stream;
  if (dt == VITB_BF16) {
    VITB_ATTN_DISPATCH(launch_bwd_bf16, qkv, d_o, lse, dqkv, B, T, heads, scale, st);
  }
  const size_t smem = ((size_t)4 * T * d + (size_t)2 * T * T) * sizeof(float);
  if (set_dyn_smem((const void*)attn_bwd_f32_kernel, smem)) return -1;
  attn_bwd_f32_kernel<<<B * heads, 128, smem, st>>>((const float*)qkv, (const float*)d_o, lse, (float*)dqkv, T, heads, d, scale);
  VITB_LAUNCH_OK();
  return 0;
}

}  // extern "C"
