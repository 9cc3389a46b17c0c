// gemm_ts.cu — EXPERIMENTAL (off unless VITB_GEMM_TS=1): resident-weight GEMM with the weight block in TENSOR MEMORY.
//
//   out[m, n] = epilogue( sum_k A[m, k] * Wv[n, k] )      A: activations (M x Kred, K-major), 128 output columns n per CTA
//
// Round-1 measurements (tools/tmemw_proto.cu, profiles/r1_tmemw_proto.log): with the 128 x Kred weight block as the M-side operand
// of tcgen05.mma in its TMEM-A form, a 128x128x384 tile takes 1717-1788 cycles (78-89 % of the MMA floor) against 3028 in the
// shared-memory-resident dual-stream kernel of gemm_tc.cu, and shared memory is free for a deeper activation ring.  The price: the
// accumulator is the TRANSPOSED tile (TMEM lane = output column n, TMEM column = activation row m), so the epilogue transposes on
// its way to the TMA-store slabs (a lane pair exchanges one value per two rows, then 32-bit shared stores).
//
//   warp 0      TMA producer: activation tiles (128 rows x 64 k, 128B swizzle) through a ring of kStages x kKps x 16 KB
//   warp 1      MMA issuer:   D[n, m] (+)= W[tmem: n, k16] * A[smem: m, k16]^T, two accumulators of 128 columns
//   warps 2-9   prologue: weight block global -> (TMA) ring -> registers -> tcgen05.st into TMEM columns [256, 256 + Kred / 2);
//               then epilogue: warp (q, h) owns output columns 32 q .. 32 q + 31 (its TMEM lane quarter) of activation rows
//               64 h .. 64 h + 63 of every tile; slab = 64 rows x 64 B, no swizzle, stored by TMA (rows >= M clipped)
// Weight layouts: fwd   Wv[n, k] = w[n][k]  (nn.Linear weight, K-major: rows straight into TMEM lanes)
//                 dgrad Wv[n, k] = w[k][n]  (the same weight read MN-major: transposed on the way through shared memory)
// Same epilogue arithmetic as gemm_tc.cu (bias, pre-activation save, exact GELU, residual; or gelu'(z) for dgrad).
// Status (end of round 1): `VITB_GEMM_TS=1 pytest tests/test_gpu_kernels.py -k "(test_gemm_fwd or test_gemm_dgrad) and 20000"` passes on
// B200 (forward plain and GELU + residual + pre-activation, dgrad with and without gelu'(z), M = 20000 not a multiple of 128);
// its speed has not been measured yet (the round's GPU budget ended) — round 2 starts with `VITB_GEMM_TS=1 python bench.py`.
#include "common.cuh"
#include "gemm_internal.h"

namespace vitb {
namespace ts {

constexpr int BM = 128, BNW = 128, BK = 64, kMaxKB = 6;
constexpr int kStages = 4, kKps = 2;
constexpr uint32_t kTile = BM * BK * 2;          // 16 KB
constexpr uint32_t kStage = kKps * kTile;        // 32 KB
constexpr uint32_t kRing = kStages * kStage;     // 128 KB (>= the 96 KB the weight block needs on its way in)
constexpr int kEpiWarps = 8;
constexpr uint32_t kSlab = 64 * 64;              // 64 rows x 32 bf16
constexpr uint32_t kSlabs = 2 * kEpiWarps * kSlab;  // out (+ input operand, consumed in place) and pre-activation
constexpr uint32_t kBarOff = kRing + kSlabs;
constexpr int kNumBars = 2 * kStages + 4 + 2 + kEpiWarps;  // full, empty, tfull[2], tempty[2], wfull, wready, in[8]
constexpr uint32_t kDynBytes = kBarOff + kNumBars * 8 + 16 + 1024;
constexpr uint32_t kColW = 256;
constexpr int kThreads = (2 + kEpiWarps) * 32;
constexpr long long kSpin = 20LL * 1000 * 1000 * 1000;
static_assert(kDynBytes <= 232448, "shared memory plan exceeds 227 KB");
static_assert(kRing >= kMaxKB * kTile, "the weight block is staged in the ring");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && clock64() - t0 > kSpin) __trap();  // a lost arrival must not hang the GPU
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(map),
               "r"(bar), "r"(c0), "r"(c1), "r"(0)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(0) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P1;\nelect.sync _|P1, 0xffffffff;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile, 128-byte swizzle: LBO unused (16 B), SBO = 1024 B, descriptor version 1, swizzle mode 2
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | ((16u >> 4) << 16);
  const uint32_t hi = (uint32_t)((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ float bf16_bits_to_f(uint32_t b) { return __uint_as_float(b << 16); }

struct Args {
  const float* bias;  // [Nout] or null (fwd)
  int M, kblocks, nblocks, members;
  int mode;           // EPI_FWD / EPI_DGRAD
  int gelu, has_in, has_pre, w_mn;
};

// slab element (row j, column `lane`) of a 64-row x 32-column bf16 slab without swizzle
__device__ __forceinline__ uint32_t slab_addr(uint32_t slab, int j, int col) { return slab + (uint32_t)j * 64u + (uint32_t)col * 2u; }

// Transposing store: this thread holds v[j] = value of (row j, column lane), j = 0..63.  Lane pairs trade one value per two rows:
// the even lane ends up with (row j: columns lane, lane + 1), the odd lane with (row j + 1: columns lane - 1, lane) -> 32-bit stores,
// rows j and j + 1 land in different bank halves.
__device__ __forceinline__ void slab_store_t(uint32_t slab, int lane, const float (&v)[64]) {
  const bool odd = lane & 1;
#pragma unroll
  for (int j = 0; j < 64; j += 2) {
    const float send = odd ? v[j] : v[j + 1];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
    const uint32_t word = odd ? pack_bf16x2(recv, v[j + 1]) : pack_bf16x2(v[j], recv);
    sts32(slab_addr(slab, odd ? j + 1 : j, lane & ~1), word);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
    gemm_ts_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out,
                   const __grid_constant__ CUtensorMap map_pre, const __grid_constant__ CUtensorMap map_in, const Args p) {
  pdl_trigger();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slabs = base + kRing;
  const uint32_t bars = base + kBarOff;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (kStages + s); };
  auto tfull = [&](int a) { return bars + 8u * (2 * kStages + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * kStages + 2 + a); };
  const uint32_t wfull = bars + 8u * (2 * kStages + 4), wready = bars + 8u * (2 * kStages + 5);
  auto in_bar = [&](int w) { return bars + 8u * (2 * kStages + 6 + w); };
  const uint32_t tmem_slot = bars + 8u * kNumBars;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), kEpiWarps); }
    mbar_init(wfull, 1);
    mbar_init(wready, kEpiWarps);
    for (int w = 0; w < kEpiWarps; ++w) mbar_init(in_bar(w), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  const int nb = blockIdx.x % p.nblocks, member = blockIdx.x / p.nblocks;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int KBn = p.kblocks;
  // D fp32, A / B bf16, both K-major; N (activation rows per tile) = 128, M (weight rows) = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BM >> 3) << 17) | ((uint32_t)(BNW >> 4) << 24);

  if (warp == 0) {
    // ================= TMA producer =================
    const bool leader = elect_one();
    if (leader) {  // the weight block, staged in the (still unused) ring
      mbar_arrive_expect_tx(wfull, (uint32_t)KBn * kTile);
      for (int kb = 0; kb < KBn; ++kb) {
        if (!p.w_mn) {
          tma_load_3d(base + kb * kTile, &map_w, wfull, kb * BK, nb * BNW);  // box: 64 k x 128 weight rows
        } else {
          tma_load_3d(base + kb * kTile, &map_w, wfull, nb * BNW, kb * BK);            // box: 64 output columns x 64 reduction rows
          tma_load_3d(base + kb * kTile + kTile / 2, &map_w, wfull, nb * BNW + 64, kb * BK);
        }
      }
    }
    __syncwarp();
    mbar_wait(wready, 0);  // the weight block has left the ring
    int stage = 0;
    uint32_t phase = 0;
    for (int mt = member; mt < m_tiles; mt += p.members) {
      for (int kb = 0; kb < KBn; kb += kKps) {
        const int nsub = min(kKps, KBn - kb);
        mbar_wait(empty(stage), phase ^ 1u);
        if (leader) {
          mbar_arrive_expect_tx(full(stage), (uint32_t)nsub * kTile);
          for (int j = 0; j < nsub; ++j) tma_load_3d(base + stage * kStage + j * kTile, &map_a, full(stage), (kb + j) * BK, mt * BM);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const bool leader = elect_one();
    mbar_wait(wready, 0);
    tc_fence_after();
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int mt = member; mt < m_tiles; mt += p.members) {
      mbar_wait(tempty(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)acc * 128u;
      for (int kb = 0; kb < KBn; kb += kKps) {
        const int nsub = min(kKps, KBn - kb);
        mbar_wait(full(stage), phase);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int j = 0; j < kKps; ++j) {
            if (j < nsub) {
#pragma unroll
              for (int kk = 0; kk < BK / 16; ++kk) {
                const uint32_t a_t = tmem_base + kColW + (uint32_t)((kb + j) * (BK / 16) + kk) * 8u;
                tc_mma_ts(d, a_t, make_desc(base + stage * kStage + j * kTile + kk * 32), idesc, (kb + j > 0 || kk > 0) ? 1u : 0u);
              }
            }
          }
          tc_commit(empty(stage));
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      if (leader) tc_commit(tfull(acc));
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ================= weight block -> TMEM, then epilogue =================
    const int ew = warp - 2, q = warp & 3, h = ew >> 2;
    const int r = q * 32 + lane;  // weight row of this thread = TMEM lane = output column inside the CTA's block
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    mbar_wait(wfull, 0);
    {
      const int steps = KBn * (BK / 16);  // k-steps of 16; the two warps of a lane quarter take half each
      const int s0 = h * (steps / 2), s1 = h == 0 ? steps / 2 : steps;
      for (int ks = s0; ks < s1; ++ks) {
        const int kb = ks >> 2, g = ks & 3;
        uint32_t v[8];
        if (!p.w_mn) {  // row r of the 128 x 64 tile: chunks 2g, 2g + 1 (16 B each) under the 128B swizzle
          const uint32_t row = base + kb * kTile + (uint32_t)r * 128u;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t a = row + ((uint32_t)((2 * g + c) ^ (r & 7)) << 4);
            v[4 * c + 0] = lds32(a); v[4 * c + 1] = lds32(a + 4); v[4 * c + 2] = lds32(a + 8); v[4 * c + 3] = lds32(a + 12);
          }
        } else {  // transposed: column r % 64 of the 64 (reduction rows) x 64 tile number r / 64, rows 16 g .. 16 g + 15
          const uint32_t tile = base + kb * kTile + (uint32_t)(r >> 6) * (kTile / 2);
          const int cc = r & 63;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int n0 = 16 * g + 2 * i, n1 = n0 + 1;
            const uint32_t lo = lds16(tile + (uint32_t)n0 * 128u + ((uint32_t)((cc >> 3) ^ (n0 & 7)) << 4) + (uint32_t)(cc & 7) * 2u);
            const uint32_t hi = lds16(tile + (uint32_t)n1 * 128u + ((uint32_t)((cc >> 3) ^ (n1 & 7)) << 4) + (uint32_t)(cc & 7) * 2u);
            v[i] = lo | (hi << 16);
          }
        }
        tmem_st8(lane_base + kColW + (uint32_t)ks * 8u, v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(wready);
    }

    const uint32_t slab_out = slabs + (uint32_t)ew * kSlab;
    const uint32_t slab_pre = slabs + (uint32_t)(kEpiWarps + ew) * kSlab;
    const uint32_t slab_in = slab_out;  // the input operand is consumed in place
    const int ncol = nb * BNW + r;
    const float bias_n = (p.mode == EPI_FWD && p.bias != nullptr) ? p.bias[ncol] : 0.f;
    int acc = 0;
    uint32_t acc_phase = 0, in_phase = 0;
    bool stores_pending = false;
    for (int mt = member; mt < m_tiles; mt += p.members) {
      const int n0 = nb * BNW + q * 32, m0 = mt * BM + h * 64;
      if (p.has_in) {
        if (lane == 0) {
          if (stores_pending) tma_store_wait_read();  // the slab doubles as the previous tile's output slab
          mbar_arrive_expect_tx(in_bar(ew), kSlab);
          tma_load_3d(slab_in, &map_in, in_bar(ew), n0, m0);
        }
      }
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      const uint32_t t = lane_base + (uint32_t)acc * 128u + (uint32_t)h * 64u;
      tmem_ld32(t, r0);
      tmem_ld32(t + 32u, r1);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(acc));  // accumulator drained into registers
      float v[64];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = __uint_as_float(r0[j]);
        v[32 + j] = __uint_as_float(r1[j]);
      }
      if (stores_pending) {
        if (lane == 0 && !p.has_in) tma_store_wait_read();
        __syncwarp();
      }
      if (p.mode == EPI_FWD) {
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] += bias_n;
        if (p.has_pre) slab_store_t(slab_pre, lane, v);
        if (p.gelu) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = gelu_f(v[j]);
        }
        if (p.has_in) {
          mbar_wait(in_bar(ew), in_phase);
          in_phase ^= 1u;
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] += bf16_bits_to_f(lds16(slab_addr(slab_in, j, lane)));
          __syncwarp();  // every lane has read its column before any lane overwrites the slab
        }
      } else {
        if (p.has_in) {
          mbar_wait(in_bar(ew), in_phase);
          in_phase ^= 1u;
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] *= gelu_grad_f(bf16_bits_to_f(lds16(slab_addr(slab_in, j, lane))));
          __syncwarp();
        }
      }
      slab_store_t(slab_out, lane, v);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&map_out, slab_out, n0, m0);  // rows >= M are clipped by the tensor map
        if (p.has_pre) tma_store_3d(&map_pre, slab_pre, n0, m0);
        tma_store_commit();
      }
      stores_pending = true;
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (stores_pending && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace ts

bool ts_gemm_ok(int M, int Nout, int Kred) {
  return Nout % ts::BNW == 0 && Kred % ts::BK == 0 && Kred / ts::BK <= ts::kMaxKB && Kred >= 2 * ts::BK && Nout / ts::BNW <= kNumSMs &&
         (M + ts::BM - 1) / ts::BM >= 2 * (kNumSMs / (Nout / ts::BNW));
}

// mode EPI_FWD:   out[M, Nout] = act(a[M, Kred] w[Nout, Kred]^T + bias) (+ in), optional pre-activation copy
// mode EPI_DGRAD: out[M, Nout] = (a[M, Kred] w[Kred, Nout]) * gelu'(in)            (w_mn_major = true)
int ts_gemm_launch(int mode, const void* a, const void* w, const float* bias, const void* in, void* out, void* pre, int M, int Nout, int Kred, int gelu,
                   bool w_mn_major, cudaStream_t st) {
  CUtensorMap ma, mw, mo, mp, mi;
  if (make_tma_map_3d_bf16(&ma, a, Kred, M, 1, (uint64_t)Kred * 2, (uint64_t)Kred * 2 * M, ts::BK, ts::BM, 1, 128)) return -1;
  if (!w_mn_major) {
    if (make_tma_map_3d_bf16(&mw, w, Kred, Nout, 1, (uint64_t)Kred * 2, (uint64_t)Kred * 2 * Nout, ts::BK, ts::BNW, 1, 128)) return -1;
  } else {  // w is [Kred rows][Nout columns]
    if (make_tma_map_3d_bf16(&mw, w, Nout, Kred, 1, (uint64_t)Nout * 2, (uint64_t)Nout * 2 * Kred, 64, 64, 1, 128)) return -1;
  }
  if (make_tma_map_3d_bf16(&mo, out, Nout, M, 1, (uint64_t)Nout * 2, (uint64_t)Nout * 2 * M, 32, 64, 1, 0)) return -1;
  mp = mo;
  mi = mo;
  if (pre && make_tma_map_3d_bf16(&mp, pre, Nout, M, 1, (uint64_t)Nout * 2, (uint64_t)Nout * 2 * M, 32, 64, 1, 0)) return -1;
  if (in && make_tma_map_3d_bf16(&mi, in, Nout, M, 1, (uint64_t)Nout * 2, (uint64_t)Nout * 2 * M, 32, 64, 1, 0)) return -1;
  ts::Args p = {};
  p.bias = bias; p.M = M; p.kblocks = Kred / ts::BK; p.nblocks = Nout / ts::BNW; p.members = kNumSMs / p.nblocks;
  p.mode = mode; p.gelu = gelu; p.has_in = in != nullptr; p.has_pre = pre != nullptr; p.w_mn = w_mn_major ? 1 : 0;
  static bool configured = false;
  if (!configured) {
    VITB_CUDA_OK(cudaFuncSetAttribute(ts::gemm_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts::kDynBytes));
    configured = true;
  }
  VITB_LAUNCH((ts::gemm_ts_kernel), p.nblocks * p.members, ts::kThreads, ts::kDynBytes, st, ma, mw, mo, mp, mi, p);
  VITB_LAUNCH_OK();
  return 0;
}

}  // namespace vitb
