// gemm_bwd_fused.cuh — backward of one nn.Linear  y = x · Wᵀ  (W: [N, K]) in ONE pass over dY (included by gemm_tc.cu):
//
//     dX [M, K] = dY [M, N] · W            ( ∘ gelu'(Z) when the Linear's input was an activation: layers.py:33-37 )
//     dW [N, K] = dYᵀ · X                  ( fp32 per-CTA partials, reduced in fixed order afterwards )
//     colsum(dX)                           ( optional: the bias gradient of the Linear that produced Z )
//
// autograd runs these as two GEMMs that both stream dY from HBM (layers.py:33-37, 85, 102 backward).  Here a CTA owns one
// 128-column block j of K and a share of the 128-row blocks of M: it keeps W[:, j] (N x 128, <= 96 KB) resident, and for every
// row block streams dY[rows, 0:N] ONCE through shared memory.  The same bytes feed both products — as the K-major A operand of
// the dgrad MMAs (reduction over n) and, reinterpreted, as the MN-major A operand of the wgrad MMAs (reduction over the rows) —
// so dY crosses HBM and the L2->SM fabric once instead of twice, X only in 128-column slices, and one launch (prologue, weight
// load, tail) replaces two plus a split-K second pass per row block.  Tensor memory holds the dX tile (128 columns) and the CTA's
// whole dW slice (N/128 accumulators of 128 columns, accumulated over all of its row blocks): 512 columns exactly at N = 384.
//
// Warps: 0 TMA producer of dY (and W), 1 dgrad MMA issuer, 2 wgrad MMA issuer (two issuing threads keep the pipe busy at this tile
// width, see gemm_tc.cu), 3 TMA producer of X, 4-11 epilogue (dX tiles through swizzled slabs and TMA stores as in gemm_tc_kernel;
// dW slice once at the end).
// Shared memory (227 KB): W block 96 KB | dY ring: 2 stages of a 128-row x 128-column pair of 64-column sub-tiles (2 x 32 KB) |
// X: 2 buffers of 128 rows x 128 columns (2 x 32 KB) | barriers.  X of a row block is live for the whole row block (every dY pair
// multiplies it), so it must be double-buffered or the wgrad issuer waits a full load latency per row block (the first version
// of this kernel, 12k cycles per row block against a 3k MMA floor); there is no room left for epilogue slabs, so the epilogue of
// row block t stages its tile in the X buffer of row block t, which the wgrad MMAs have just finished reading, and hands the
// buffer to the X producer (for row block t + 2) when its TMA store has read it.
#pragma once

namespace vitb {

struct BwArgs {
  int M, N, K;
  int num_m_blocks, members;  // 128-row blocks of M; CTAs per 128-column block of K (each walks row blocks member, member + members, ...)
  int has_in;                 // tma_in valid: z for gelu'
  float* dw_part;             // [members][N][K] fp32
  float* csum_part;           // [members * 4][K] fp32 column sums of the dX rows this CTA produced, or null
  int pf_tiles;               // L2 prefetch distance in row blocks (0 = off)
};

template <int NB>  // N / 128
struct BwSmem {
  static constexpr uint32_t kSub = 128 * 64 * 2;               // one 128-row x 64-column bf16 sub-tile (128B swizzle)
  static constexpr uint32_t kWBytes = 2 * NB * kSub;           // N/64 reduction blocks of [64 n-rows x 128 k-columns]
  static constexpr int kDyStages = 2, kXBufs = 2;
  static constexpr uint32_t kDyStage = 2 * kSub;               // a pair: 128 rows x 128 columns of dY
  static constexpr uint32_t kXHalf = kSub;                     // 64 rows x 128 columns of X (two 64-column atoms of 64 rows)
  static constexpr uint32_t kXBuf = 2 * kXHalf;                // X of one row block; afterwards the 8 epilogue slabs of that row block
  static_assert(kXBuf == kEpiWarps * kSlabBytes, "the epilogue slabs alias one X buffer exactly");
  static constexpr uint32_t kDyOff = kWBytes;
  static constexpr uint32_t kXOff = kDyOff + kDyStages * kDyStage;
  static constexpr uint32_t kBarOff = kXOff + kXBufs * kXBuf;
  // dy_full[2] dy_empty[2] xb_full[2] xb_gdone[2] xb_free[2] tfull tempty wfull d2full in[8]
  static constexpr uint32_t kNumBars = 2 * kDyStages + 3 * kXBufs + 4 + kEpiWarps;
  static constexpr uint32_t kTotal = kBarOff + kNumBars * 8 + 16;
  static constexpr uint32_t kDynBytes = kTotal + 1024;
};

template <int NB>
__global__ void __launch_bounds__(384, 1)
    gemm_bwd_fused_kernel(const __grid_constant__ CUtensorMap tma_dy, const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w,
                          const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_in, const BwArgs p) {
  using S = BwSmem<NB>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();

  auto w_res = [&](int kb) { return smem_base + (uint32_t)kb * S::kSub; };          // reduction block kb of the resident W slice
  auto dy_stage = [&](int s) { return smem_base + S::kDyOff + (uint32_t)s * S::kDyStage; };
  auto x_buf = [&](int b) { return smem_base + S::kXOff + (uint32_t)b * S::kXBuf; };
  const uint32_t bar_base = smem_base + S::kBarOff;
  auto dy_full = [&](int s) { return bar_base + (uint32_t)s * 8; };
  auto dy_empty = [&](int s) { return bar_base + (uint32_t)(S::kDyStages + s) * 8; };
  auto xb_full = [&](int b) { return bar_base + (uint32_t)(2 * S::kDyStages + b) * 8; };                  // X of a row block has landed
  auto xb_gdone = [&](int b) { return bar_base + (uint32_t)(2 * S::kDyStages + S::kXBufs + b) * 8; };     // the wgrad MMAs have read it
  auto xb_free = [&](int b) { return bar_base + (uint32_t)(2 * S::kDyStages + 2 * S::kXBufs + b) * 8; };  // the epilogue is done with it
  const uint32_t misc = bar_base + (uint32_t)(2 * S::kDyStages + 3 * S::kXBufs) * 8;
  const uint32_t tfull_bar = misc, tempty_bar = misc + 8, wfull_bar = misc + 16, d2full_bar = misc + 24;
  auto in_bar = [&](int w) { return misc + 32 + (uint32_t)w * 8; };
  const uint32_t tmem_slot = bar_base + S::kNumBars * 8;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + S::kBarOff + S::kNumBars * 8);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tma_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tma_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tma_w) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tma_out) : "memory");
    for (int s = 0; s < S::kDyStages; ++s) {
      mbar_init(dy_full(s), 1);
      mbar_init(dy_empty(s), 2);  // one tcgen05.commit of each issuing warp
    }
    for (int b = 0; b < S::kXBufs; ++b) {
      mbar_init(xb_full(b), 1);
      mbar_init(xb_gdone(b), 1);
      mbar_init(xb_free(b), kEpiWarps);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, kEpiWarps);
    mbar_init(wfull_bar, 1);
    mbar_init(d2full_bar, 1);
    for (int w = 0; w < kEpiWarps; ++w) mbar_init(in_bar(w), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();  // operands may belong to the preceding kernel

  const int nblocks = p.K / 128;
  const int jblock = (int)blockIdx.x % nblocks;   // this CTA's 128-column block of K
  const int member = (int)blockIdx.x / nblocks;   // < p.members
  const int my_blocks = (p.num_m_blocks - member + p.members - 1) / p.members;  // >= 1 by construction of the grid
  // TMEM: dX tile at columns [0, 128), dW accumulator of row block q of N at [128 (q + 1), 128 (q + 2))
  constexpr uint32_t kLboDy = 2 * 8192;  // distance between the two 64-column atoms of a dY pair (128-row sub-tiles)
  constexpr uint32_t kLbo64 = 8192;      // ... of 64-row tiles (W reduction blocks, X stages)
  constexpr uint32_t kStepK = (UK * 2) >> 4, kStepMN = (UK * 128) >> 4;  // descriptor advance per 16 reduction elements

  if (warp == 0) {
    // ================= TMA producer =================
    const bool leader = elect_one();
    if (leader) {
      mbar_arrive_expect_tx(wfull_bar, S::kWBytes);
      for (int kb = 0; kb < 2 * NB; ++kb) {
        tma_load_2d(w_res(kb), &tma_w, wfull_bar, jblock * 128, kb * 64);
        tma_load_2d(w_res(kb) + kLbo64, &tma_w, wfull_bar, jblock * 128 + 64, kb * 64);
      }
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < my_blocks; ++it) {
      const int m0 = (member + it * p.members) * BM;
      if (leader && p.pf_tiles > 0 && it + p.pf_tiles < my_blocks) {
        // ask L2 for a row block further ahead than the rings reach; the CTAs that share the row block (one per column block)
        // split its 64-column pieces of dY between them
        const int pm0 = (member + (it + p.pf_tiles) * p.members) * BM;
        for (int kb = jblock; kb < 2 * NB; kb += nblocks) tma_prefetch_2d(&tma_dy, kb * 64, pm0);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tma_prefetch_2d(&tma_x, jblock * 128 + 64 * c, pm0);
          tma_prefetch_2d(&tma_x, jblock * 128 + 64 * c, pm0 + 64);
          if (p.has_in) {
#pragma unroll
            for (int r = 0; r < 4; ++r) tma_prefetch_2d(&tma_in, jblock * 128 + 64 * c, pm0 + 32 * r);
          }
        }
      }
#pragma unroll 1
      for (int q = 0; q < NB; ++q) {
        mbar_wait(dy_empty(stage), phase ^ 1u);
        if (leader) {
          mbar_arrive_expect_tx(dy_full(stage), S::kDyStage);
          tma_load_2d(dy_stage(stage), &tma_dy, dy_full(stage), (2 * q) * 64, m0);
          tma_load_2d(dy_stage(stage) + S::kSub, &tma_dy, dy_full(stage), (2 * q + 1) * 64, m0);
        }
        __syncwarp();
        if (++stage == S::kDyStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 3) {
    // ================= TMA producer of X: row block `it` goes to buffer it & 1 once the epilogue of row block it - 2 has
    // released it (its own producer so that a late release never holds up the dY ring) =================
    const bool leader = elect_one();
    for (int it = 0; it < my_blocks; ++it) {
      const int m0 = (member + it * p.members) * BM;
      const int b = it & 1;
      mbar_wait(xb_free(b), ((uint32_t)(it >> 1) & 1u) ^ 1u);
      if (leader) {
        mbar_arrive_expect_tx(xb_full(b), S::kXBuf);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          tma_load_2d(x_buf(b) + (uint32_t)h * S::kXHalf, &tma_x, xb_full(b), jblock * 128, m0 + 64 * h);
          tma_load_2d(x_buf(b) + (uint32_t)h * S::kXHalf + kLbo64, &tma_x, xb_full(b), jblock * 128 + 64, m0 + 64 * h);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= dgrad issuer: dX tile += dY pair (K-major A) · resident W blocks (MN-major B) =================
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t a_lo0 = desc_lo(dy_stage(0), 16), b_lo0 = desc_lo(w_res(0), kLbo64);
    mbar_wait(wfull_bar, 0);
    tc_fence_after();
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < my_blocks; ++it) {
      mbar_wait(tempty_bar, (uint32_t)(it & 1) ^ 1u);  // the epilogue has drained the previous dX tile
      tc_fence_after();
#pragma unroll 1
      for (int q = 0; q < NB; ++q) {
        mbar_wait(dy_full(stage), phase);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const uint32_t a_lo = a_lo0 + (uint32_t)stage * (S::kDyStage >> 4) + (uint32_t)sub * (S::kSub >> 4);
            const uint32_t b_lo = b_lo0 + (uint32_t)(2 * q + sub) * (S::kSub >> 4);
#pragma unroll
            for (int kk = 0; kk < BK / UK; ++kk)
              tc_mma_bf16(tmem_base, desc_from_lo(a_lo + kk * kStepK), desc_from_lo(b_lo + kk * kStepMN), idesc, (q | sub | kk) != 0 ? 1u : 0u);
          }
          tc_commit(dy_empty(stage));
          if (q == NB - 1) tc_commit(tfull_bar);
        }
        __syncwarp();
        if (++stage == S::kDyStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 2) {
    // ================= wgrad issuer: dW[n block q, this k block] += (dY pair)ᵀ (MN-major A, reduction over the 128 rows)
    //                   · X rows (MN-major B), in two halves of 64 rows =================
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t a_lo0 = desc_lo(dy_stage(0), kLboDy), b_lo0 = desc_lo(x_buf(0), kLbo64);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < my_blocks; ++it) {
      const int xb = it & 1;
#pragma unroll 1
      for (int q = 0; q < NB; ++q) {
        mbar_wait(dy_full(stage), phase);
        if (q == 0) mbar_wait(xb_full(xb), (uint32_t)(it >> 1) & 1u);
        tc_fence_after();
        if (leader) {
          const uint32_t d2 = tmem_base + 128u * (uint32_t)(q + 1);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t a_lo = a_lo0 + (uint32_t)stage * (S::kDyStage >> 4) + (uint32_t)h * (8192u >> 4);  // rows 64 h .. of both atoms
            const uint32_t b_lo = b_lo0 + (uint32_t)xb * (S::kXBuf >> 4) + (uint32_t)h * (S::kXHalf >> 4);
#pragma unroll
            for (int kk = 0; kk < 64 / UK; ++kk)
              tc_mma_bf16(d2, desc_from_lo(a_lo + kk * kStepMN), desc_from_lo(b_lo + kk * kStepMN), idesc, (it | h | kk) != 0 ? 1u : 0u);
          }
          tc_commit(dy_empty(stage));
          if (q == NB - 1) tc_commit(xb_gdone(xb));  // last use of this row block's X: the buffer becomes the epilogue's staging area
        }
        __syncwarp();
        if (++stage == S::kDyStages) { stage = 0; phase ^= 1u; }
      }
    }
    if (leader) tc_commit(d2full_bar);  // the CTA's dW slice is complete
    __syncwarp();
  } else if (warp >= 4) {
    // ================= epilogue: TMEM lane quarter = warp % 4, column half = ew / 4 =================
    const int ew = warp - 4;
    const int quarter = warp & 3, half = ew >> 2;
    const int n0 = jblock * 128 + half * 64;
    uint32_t in_phase = 0;
    float cs0 = 0.f, cs1 = 0.f;  // column sums of columns n0 + 2 lane, + 1 over this warp's rows of every tile
    for (int it = 0; it < my_blocks; ++it) {
      const int m0 = (member + it * p.members) * BM;
      const int xb = it & 1;
      const uint32_t slab = x_buf(xb) + (uint32_t)ew * kSlabBytes;  // z in, dX out — once the wgrad MMAs have read X from this buffer
      mbar_wait(tfull_bar, (uint32_t)(it & 1));
      tc_fence_after();
      uint32_t raw0[32], raw1[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 64);
      tc_ld32(taddr, raw0);
      tc_ld32(taddr + 32u, raw1);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);  // the dX accumulator is in registers: the dgrad issuer may start the next row block
      float v[64];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = __uint_as_float(raw0[j]);
        v[32 + j] = __uint_as_float(raw1[j]);
      }
      mbar_wait(xb_gdone(xb), (uint32_t)(it >> 1) & 1u);
      if (p.has_in) {
        if (lane == 0) {
          mbar_arrive_expect_tx(in_bar(ew), kSlabBytes);
          tma_load_2d(slab, &tma_in, in_bar(ew), n0, m0 + quarter * 32);
        }
        mbar_wait(in_bar(ew), in_phase);
        in_phase ^= 1u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float z[8];
          slab_load_chunk8(slab, lane, j, z);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[8 * j + i] *= gelu_grad_bf16_f(z[i]);
        }
      }
      slab_store_row64(slab, lane, v);  // (a lane reads and writes only its own row of the slab)
      __syncwarp();
      if (p.csum_part != nullptr) {
        // column sums of the bf16-rounded tile (what a later reader of dX sees): lane l owns columns 2l, 2l+1; rows beyond M are
        // exact zeros (zero-filled operands)
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          const uint32_t addr = slab + (uint32_t)r * 128u + ((((uint32_t)lane >> 2) ^ (uint32_t)(r & 7)) << 4) + (((uint32_t)lane & 3u) << 2);
          uint32_t u;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(addr) : "memory");
          const float2 f = unpack_bf16x2(u);
          cs0 += f.x;
          cs1 += f.y;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tma_out, slab, n0, m0 + quarter * 32);  // rows >= M are clipped by the tensor map
        tma_store_commit();
        tma_store_wait_read();         // the store has read the slab:
        mbar_arrive(xb_free(xb));      // the buffer may receive X of row block it + 2
      }
      __syncwarp();
    }
    if (p.csum_part != nullptr) {
      float* o = p.csum_part + (size_t)(member * 4 + quarter) * p.K + n0 + 2 * lane;
      *reinterpret_cast<float2*>(o) = make_float2(cs0, cs1);
    }
    // the CTA's dW slice: row block q of N x this CTA's 128 columns, fp32, straight from registers (256 contiguous bytes per thread)
    mbar_wait(d2full_bar, 0);
    tc_fence_after();
    float* part = p.dw_part + (size_t)member * p.N * p.K;
#pragma unroll 1
    for (int q = 0; q < NB; ++q) {
      float* o = part + (size_t)(q * 128 + quarter * 32 + lane) * p.K + n0;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t raw[32];
        tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + 128u * (uint32_t)(q + 1) + (uint32_t)(half * 64 + 32 * c), raw);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          reinterpret_cast<float4*>(o + 32 * c)[j] = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                                                 __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace vitb
