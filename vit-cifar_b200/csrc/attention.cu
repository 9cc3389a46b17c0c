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
__device__ __forceinline__ void stsm_x4(void* p, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
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

// One CTA per (image, head); low register count (scores are processed in 16-key chunks) so that 4 CTAs share an SM.
//   D_i = sum_d dO_id * O_id  (= rowsum(P ∘ dP), flash-attention identity) from the saved forward output, so a chunk's
//   dS can be formed without holding the whole score row.
template <int D, int NT16>
__global__ void __launch_bounds__(32 * NT16, (D == 32 ? (NT16 <= 2 ? 8 : (NT16 <= 5 ? 4 : 1)) : (NT16 <= 2 ? 4 : (NT16 <= 5 ? 2 : 1))))
    attn_bwd_bf16_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                         const float* __restrict__ lse, bf16* __restrict__ dqkv, int T, int heads, float scale) {
  constexpr int TP = 16 * NT16, LD = D + 8, LP = TP + 8, TILE = TP * LD, CH = D / 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_attn);
  bf16* sK = sQ + TILE;
  bf16* sV = sK + TILE;
  bf16* sdO = sV + TILE;
  bf16* sP = sdO + TILE;     // [TP][LP]
  bf16* sdS = sP + TP * LP;  // [TP][LP]
  float* sD = reinterpret_cast<float*>(sdS + TP * LP);  // [TP]
  float* sL = sD + TP;                                   // [TP] lse * log2(e)
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int Hd = heads * D;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x;
  const bf16* base = qkv + (int64_t)b * T * 3 * Hd + h * D;
  const int64_t orow = (int64_t)b * T * Hd + h * D;
  // all tile loads are asynchronous (one round trip); the other CTAs resident on the SM compute meanwhile
  load_head_tile_async<D, TP>(sQ, base, 3 * Hd, T, tid, nthr);
  load_head_tile_async<D, TP>(sK, base + Hd, 3 * Hd, T, tid, nthr);
  load_head_tile_async<D, TP>(sV, base + 2 * Hd, 3 * Hd, T, tid, nthr);
  load_head_tile_async<D, TP>(sdO, d_o + orow, Hd, T, tid, nthr);
  cp_async_commit();
  zero_pad_rows<D, TP>(sQ, T, tid, nthr);
  zero_pad_rows<D, TP>(sK, T, tid, nthr);
  zero_pad_rows<D, TP>(sV, T, tid, nthr);
  zero_pad_rows<D, TP>(sdO, T, tid, nthr);
  for (int r = tid; r < TP; r += nthr) sL[r] = r < T ? lse[((int64_t)b * heads + h) * T + r] * kLog2e : 0.f;
  // D_i = sum_d dO_id * O_id straight from global memory: CH threads per row, 8 elements each
  for (int idx = tid; idx < TP * CH; idx += nthr) {
    const int r = idx / CH, c = idx % CH;
    float v = 0.f;
    if (r < T) {
      const uint4 a = *reinterpret_cast<const uint4*>(d_o + orow + (int64_t)r * Hd + c * 8);
      const uint4 q = *reinterpret_cast<const uint4*>(o + orow + (int64_t)r * Hd + c * 8);
      const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
      const float2 q0 = unpack_bf16x2(q.x), q1 = unpack_bf16x2(q.y), q2 = unpack_bf16x2(q.z), q3 = unpack_bf16x2(q.w);
      v = (a0.x * q0.x + a0.y * q0.y) + (a1.x * q1.x + a1.y * q1.y) + (a2.x * q2.x + a2.y * q2.y) + (a3.x * q3.x + a3.y * q3.y);
    }
#pragma unroll
    for (int off = CH / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);  // CH (4 or 8) adjacent lanes share a row
    if (c == 0) sD[r] = v;
  }
  cp_async_wait<0>();
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const int r0 = 16 * warp + g, r1 = r0 + 8;
  const float sl2 = scale * kLog2e;
  const float l0 = sL[r0], l1 = sL[r1];
  const float d0 = sD[r0], d1 = sD[r1];
  const bool rv0 = r0 < T, rv1 = r1 < T;
  uint32_t aq[D / 16][4], ado[D / 16][4];
#pragma unroll
  for (int kk = 0; kk < D / 16; ++kk) {
    ldsm_x4(aq[kk], sQ + (16 * warp + (lane & 15)) * LD + kk * 16 + (lane >> 4) * 8);
    ldsm_x4(ado[kk], sdO + (16 * warp + (lane & 15)) * LD + kk * 16 + (lane >> 4) * 8);
  }
  float dq[D / 8][4];
#pragma unroll
  for (int jn = 0; jn < D / 8; ++jn) dq[jn][0] = dq[jn][1] = dq[jn][2] = dq[jn][3] = 0.f;
#pragma unroll
  for (int jc = 0; jc < NT16; ++jc) {  // 16 keys per chunk
    float s[2][4], dp[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = 2 * jc + u;
      s[u][0] = s[u][1] = s[u][2] = s[u][3] = 0.f;
      dp[u][0] = dp[u][1] = dp[u][2] = dp[u][3] = 0.f;
#pragma unroll
      for (int k2 = 0; k2 < D / 32; ++k2) {
        uint32_t bk[4], bv[4];
        ldsm_x4(bk, sK + (8 * j + (lane & 7)) * LD + k2 * 32 + (lane >> 3) * 8);
        ldsm_x4(bv, sV + (8 * j + (lane & 7)) * LD + k2 * 32 + (lane >> 3) * 8);
        mma_bf16(s[u], aq[2 * k2], bk[0], bk[1]);
        mma_bf16(s[u], aq[2 * k2 + 1], bk[2], bk[3]);
        mma_bf16(dp[u], ado[2 * k2], bv[0], bv[1]);
        mma_bf16(dp[u], ado[2 * k2 + 1], bv[2], bv[3]);
      }
      const int c = 8 * j + 2 * t;
      const bool v0 = c < T, v1 = c + 1 < T;
      s[u][0] = (v0 && rv0) ? exp2f(s[u][0] * sl2 - l0) : 0.f;
      s[u][1] = (v1 && rv0) ? exp2f(s[u][1] * sl2 - l0) : 0.f;
      s[u][2] = (v0 && rv1) ? exp2f(s[u][2] * sl2 - l1) : 0.f;
      s[u][3] = (v1 && rv1) ? exp2f(s[u][3] * sl2 - l1) : 0.f;
      // dS = P ∘ (dP − D) / sqrt(features)
      dp[u][0] = s[u][0] * (dp[u][0] - d0) * scale;
      dp[u][1] = s[u][1] * (dp[u][1] - d0) * scale;
      dp[u][2] = s[u][2] * (dp[u][2] - d1) * scale;
      dp[u][3] = s[u][3] * (dp[u][3] - d1) * scale;
    }
    uint32_t pa[4], da[4];  // A-operand fragments of this 16 x 16 chunk (also exactly what stmatrix stores)
    pa[0] = pack_bf16x2(s[0][0], s[0][1]); pa[1] = pack_bf16x2(s[0][2], s[0][3]);
    pa[2] = pack_bf16x2(s[1][0], s[1][1]); pa[3] = pack_bf16x2(s[1][2], s[1][3]);
    da[0] = pack_bf16x2(dp[0][0], dp[0][1]); da[1] = pack_bf16x2(dp[0][2], dp[0][3]);
    da[2] = pack_bf16x2(dp[1][0], dp[1][1]); da[3] = pack_bf16x2(dp[1][2], dp[1][3]);
    // matrices: (rows 0-7, keys 0-7), (rows 8-15, keys 0-7), (rows 0-7, keys 8-15), (rows 8-15, keys 8-15) of the chunk
    const int srow = 16 * warp + ((lane >> 3) & 1) * 8 + (lane & 7), scol = 16 * jc + (lane >> 4) * 8;
    stsm_x4(sP + srow * LP + scol, pa[0], pa[1], pa[2], pa[3]);
    stsm_x4(sdS + srow * LP + scol, da[0], da[1], da[2], da[3]);
    // dQ += dS_chunk · K_chunk  (K rows = reduction index -> ldmatrix.trans)
#pragma unroll
    for (int jp = 0; jp < D / 16; ++jp) {
      uint32_t bb[4];
      ldsm_x4_t(bb, sK + (16 * jc + (lane & 7) + ((lane >> 3) & 1) * 8) * LD + jp * 16 + (lane >> 4) * 8);
      mma_bf16(dq[2 * jp], da, bb[0], bb[1]);
      mma_bf16(dq[2 * jp + 1], da, bb[2], bb[3]);
    }
  }
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
  __syncthreads();  // everyone is done reading sQ/sK/sV/sdO: reuse own rows as staging

  const int rows_valid = min(16, T - 16 * warp);
  if (rows_valid > 0) {
    bf16* gd = dqkv + ((int64_t)b * T + 16 * warp) * 3 * Hd + h * D;
    store_rows16<D>(dq, sQ + 16 * warp * LD, gd, 3 * Hd, rows_valid, lane);
    store_rows16<D>(dk, sK + 16 * warp * LD, gd + Hd, 3 * Hd, rows_valid, lane);
    store_rows16<D>(dv, sV + 16 * warp * LD, gd + 2 * Hd, 3 * Hd, rows_valid, lane);
  }
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
static int launch_bwd_bf16(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int B, int T, int heads, float scale,
                           cudaStream_t st) {
  constexpr int TP = 16 * NT16, LD = D + 8, LP = TP + 8, TILE = TP * LD;
  constexpr size_t smem = ((size_t)4 * TILE + (size_t)2 * TP * LP) * sizeof(bf16) + (size_t)2 * TP * sizeof(float);
  auto kern = attn_bwd_bf16_kernel<D, NT16>;
  static bool configured = false;
  if (!configured) {
    VITB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  kern<<<B * heads, 32 * NT16, smem, st>>>((const bf16*)qkv, (const bf16*)o, (const bf16*)d_o, lse, (bf16*)dqkv, T, heads, scale);
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

int vitb_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int B, int T, int heads, int d, float scale,
                  int dt, void* stream) {
  VITB_REQUIRE(qkv && o && d_o && lse && dqkv, "attn_bwd: null pointer");
  VITB_REQUIRE(B > 0 && T > 0 && T <= 128 && heads > 0 && (d == 32 || d == 64), "attn_bwd: unsupported shape B=%d T=%d heads=%d d=%d (T<=128, d in {32,64})", B, T, heads, d);
  cudaStream_t st = (cudaStream_t)stream;
  if (dt == VITB_BF16) {
    VITB_ATTN_DISPATCH(launch_bwd_bf16, qkv, o, d_o, lse, dqkv, B, T, heads, scale, st);
  }
  const size_t smem = ((size_t)4 * T * d + (size_t)2 * T * T) * sizeof(float);
  if (set_dyn_smem((const void*)attn_bwd_f32_kernel, smem)) return -1;
  attn_bwd_f32_kernel<<<B * heads, 128, smem, st>>>((const float*)qkv, (const float*)d_o, lse, (float*)dqkv, T, heads, d, scale);
  VITB_LAUNCH_OK();
  return 0;
}

}  // extern "C"
