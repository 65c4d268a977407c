// passt_b200 — 2-CTA (tcgen05 cta_group::2) variant of the GEMM family (sm_100a only).
//
// Same math, operand layouts and epilogues as gemm.cu, but a cluster of two CTAs (one TPC) cooperates on a
// 256 (M) x 256 (N) tile: each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 N rows),
// the leader CTA issues tcgen05.mma.cta_group::2 (M=256), and each CTA's tensor core writes its own 128 rows of D
// into its own TMEM.  Per CTA the shared-memory traffic per k-block drops from 48 KB in + 48 KB out (1-CTA kernel:
// measured 68 % tensor-pipe, shared-memory-bandwidth bound) to 32 KB in + 32 KB out.
//
// Barrier protocol (all barriers live at the same smem offset in both CTAs):
//   full[s]   : leader's copy only is used.  count 1 (leader's expect_tx); BOTH CTAs' TMA loads complete_tx on it
//               (cp.async.bulk.tensor .cta_group::2 with the peer bit of the barrier address cleared).
//   empty[s]  : per CTA, count 1; tcgen05.commit ... multicast::cluster arrives on both CTAs' copies.
//   tfull[a]  : per CTA, count 1; multicast commit after the last k-block of a tile.
//   tempty[a] : leader's copy only, count 2*kEpiWarps; epilogue warps of both CTAs arrive (remote for rank 1).
#include "gemm_defs.cuh"
#include <cstdio>

namespace pb {

constexpr int kStages2 = 5;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same object in the even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2cta_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(uint16_t(3))
      : "memory");
}
// arrive on the leader CTA's copy of `bar` (works from either CTA of the pair).  Relaxed: the only thing handed over
// is tensor memory whose reads have completed (tcgen05.wait::ld + fence::before_thread_sync); a .release here costs a
// MEMBAR.ALL.GPU per tile per warp that also waits for the epilogue's outstanding bulk stores (profiles/r1_ncu_gemm_epi)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, 0;\n"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_holder) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

template <int MODE, bool BMN>
struct Gemm2Cfg {
  static constexpr int BN = 256;
  static constexpr bool kWgrad = (MODE == kWgradF32);
  static constexpr bool kAMn = kWgrad;               // A is [K, M] row-major (MN-major operand)
  static constexpr bool kBMn = kWgrad || BMN;        // B is [K, N] row-major (MN-major operand)
  static constexpr bool kOutF32 = (MODE == kRowTabF32 || MODE == kWgradF32);
  static constexpr int kABytes = BM * BK * 2;           // this CTA's 128 A rows
  static constexpr int kBBytes = (BN / 2) * BK * 2;     // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiBufBytes = 32 * 64;
  static constexpr int kEpiBytes = kEpiWarps * 2 * kEpiBufBytes;
  // after the staging buffers: 256 B pipeline barriers | 256 B per-warp aux barriers | 8 KB column-sum partials |
  // 128 B per epilogue warp for the bias values of its current block | pad
  static constexpr int kSmemBytes = kStages2 * kStageBytes + kEpiBytes + 256 + 256 + 8192 + kEpiWarps * 128 + 1024;
  static constexpr int kColsPerChunk = kOutF32 ? 16 : 32;
  static constexpr uint32_t kTmemCols = 2 * BN;
};

template <int MODE, bool BMN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, const GemmParams p) {
  using Cfg = Gemm2Cfg<MODE, BMN>;
  constexpr int BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  // both CTAs see the same dynamic-smem base offset, so the aligned offset is identical in the pair
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + kStages2 * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Cfg::kEpiBytes);
  uint64_t* full_bar = bars;                  // [kStages2]
  uint64_t* empty_bar = bars + kStages2;      // [kStages2]
  uint64_t* tfull_bar = bars + 2 * kStages2;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* aux_bar = bars + 32;              // [kEpiWarps][2] (mode 3: aux block landed in the warp's staging buffer)
  float* s_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [2 parity][4 quadrants][BN] (mode 3)
  float* s_biasw = s_part + 2 * 4 * BN;       // [kEpiWarps][32] (modes 0/1: bias of the block a warp is working on)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (MODE == kBiasGeluBf16 || MODE == kGeluGradBf16 || MODE == kRowDotBf16) tma_prefetch_desc(&tmC2);
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * kEpiWarps);
    }
    if (MODE == kGeluGradBf16 || MODE == kRowDotBf16)
      for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&aux_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta<Cfg::kTmemCols>(tmem_holder);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_gate();   // everything above is on-chip; global memory is first touched below

  // work items are 256x256 cluster tiles
  const int num_tiles = p.m_tiles * p.n_tiles * p.splits;    // m_tiles counts 256-row blocks here
  const int kb_per_split = (p.k_blocks + p.splits - 1) / p.splits;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int split = t / (p.m_tiles * p.n_tiles);
        const int tt = t - split * (p.m_tiles * p.n_tiles);
        const int m_blk = tt / p.n_tiles, n_blk = tt - m_blk * p.n_tiles;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + kb_per_split);
        const int m0 = m_blk * 256 + int(rank) * BM;          // this CTA's A rows
        const int n0 = n_blk * BN + int(rank) * (BN / 2);     // this CTA's half of B
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);   // bytes of both CTAs
          if (!Cfg::kAMn) {
            tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * BK, m0);
          } else {
#pragma unroll
            for (int g = 0; g < BM / 64; ++g)     // MN-major: boxes of [64 k rows][64 MN elements]
              tma_load_2d_2sm(sa + g * 8192, &tmA, &full_bar[stage], m0 + g * 64, kb * BK);
          }
          if (!Cfg::kBMn) {
            tma_load_2d_2sm(sb, &tmB, &full_bar[stage], kb * BK, n0);
          } else {
#pragma unroll
            for (int g = 0; g < (BN / 2) / 64; ++g)
              tma_load_2d_2sm(sb + g * 8192, &tmB, &full_bar[stage], n0 + g * 64, kb * BK);
          }
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, Cfg::kAMn ? 1 : 0, Cfg::kBMn ? 1 : 0);
      const uint64_t desc_a0 = make_smem_desc_sw128(smem_u32(smem), p.lbo_a, p.sbo_a);
      const uint64_t desc_b0 = make_smem_desc_sw128(smem_u32(smem) + Cfg::kABytes, p.lbo_b, p.sbo_b);
      const uint32_t kstep_a16 = p.kstep_a >> 4, kstep_b16 = p.kstep_b >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const int split = t / (p.m_tiles * p.n_tiles);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + kb_per_split);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          // descriptors: warp-uniform base (per stage) + per-k offset in 16-byte units; an elected lane only issues
          const uint64_t da0 = desc_a0 + uint64_t(uint32_t(stage) * (Cfg::kStageBytes >> 4));
          const uint64_t db0 = desc_b0 + uint64_t(uint32_t(stage) * (Cfg::kStageBytes >> 4));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ss_2cta(tmem_d, da0 + uint64_t(k * kstep_a16), db0 + uint64_t(k * kstep_b16), idesc,
                                (kb > kb0 || k > 0) ? 1u : 0u);
            tc_commit_2cta_mc(&empty_bar[stage]);
            if (kb == kb1 - 1) tc_commit_2cta_mc(&tfull_bar[as]);
          }
          __syncwarp();
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
        if (kb1 <= kb0 && lane == 0) tc_commit_2cta_mc(&tfull_bar[as]);
        __syncwarp();
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs; each drains its own 128 rows) =====================
    const int ew = warp - 2;
    const int q = warp & 3;
    const int slot = ew >> 2;
    uint8_t* my_epi = epi_smem + ew * 2 * Cfg::kEpiBufBytes;
    const uint32_t swz = uint32_t((lane >> 1) & 3);
    int as = 0;
    uint32_t aphase = 0;
    int buf = 0;
    int tile_parity = 0;
    uint32_t aux_phase = 0;   // bit b: parity the next wait on this warp's aux barrier b expects
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const int split = t / (p.m_tiles * p.n_tiles);
      const int tt = t - split * (p.m_tiles * p.n_tiles);
      const int m_blk = tt / p.n_tiles, n_blk = tt - m_blk * p.n_tiles;
      const int row0 = m_blk * 256 + int(rank) * BM + q * 32;
      const int row = row0 + lane;
      // mode 3: the gelu' operand (tmC2 is its tensor map) does not depend on the accumulator.  Each warp pulls its
      // 32x32 blocks with TMA into its own two staging buffers *before* waiting for the tile, reads them back
      // row-per-lane (same 64 B swizzle as the output), and then reuses the buffer for the output block.
      if (MODE == kGeluGradBf16 || MODE == kRowDotBf16) {
        if (lane == 0) {
          tma_store_wait_read<0>();          // both staging buffers have been read by the previous tile's stores
#pragma unroll
          for (int ci = 0; ci < 2; ++ci) {
            const int c0p = (slot + ci * kEpiSlots) * Cfg::kColsPerChunk;
            mbar_arrive_expect_tx(&aux_bar[ew * 2 + ci], Cfg::kEpiBufBytes);
            tma_load_2d(my_epi + ci * Cfg::kEpiBufBytes, &tmC2, &aux_bar[ew * 2 + ci], n_blk * BN + c0p, row0);
          }
        }
        __syncwarp();
      }
      // modes 0/1: lane i fetches bias[col0 + i] of the block one step ahead (here: the tile's first block, before the
      // wait), so that the L2 latency never sits between the accumulator load and its first use
      float bias_next = 0.f;
      const bool has_bias = (MODE == kBiasBf16 || MODE == kBiasGeluBf16) && p.bias != nullptr;
      if (has_bias) bias_next = __ldg(p.bias + n_blk * BN + slot * Cfg::kColsPerChunk + lane);
      // mode 2: the additive operand (token table or residual stream, + bias) of a chunk does not depend on the
      // accumulator: it is fetched one chunk ahead into registers -- the first chunk's before the wait for the tile --
      // so that its L2 latency hides behind the previous chunk's TMEM read / staging / store
      float4 add_next[4];
      const float* add_row = nullptr;
      if (MODE == kRowTabF32) {
        if (p.aux != nullptr && row < p.M)
          add_row = reinterpret_cast<const float*>(p.aux) + size_t(row % p.aux_period) * p.ld_aux + n_blk * BN;
        const int c0f = slot * Cfg::kColsPerChunk;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          if (add_row != nullptr) a = __ldg(reinterpret_cast<const float4*>(add_row + c0f) + i);
          if (p.bias != nullptr) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n_blk * BN + c0f) + i);
            a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w;
          }
          add_next[i] = a;
        }
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * BN;

#pragma unroll 1
      for (int ci = 0; ci < (BN / Cfg::kColsPerChunk + kEpiSlots - 1) / kEpiSlots; ++ci) {
        const int c0 = (slot + ci * kEpiSlots) * Cfg::kColsPerChunk;
        if (c0 >= BN) break;
        const int col0 = n_blk * BN + c0;
        if (!Cfg::kOutF32) {
          uint32_t ra[32];
          tmem_ld_x32(taddr + c0, ra);
          if (has_bias) {
            // broadcast this block's 32 bias values through the warp's smem slot; prefetch the next block's
            s_biasw[ew * 32 + lane] = bias_next;
            const int c0n = c0 + kEpiSlots * Cfg::kColsPerChunk;
            if (c0n < BN) bias_next = __ldg(p.bias + n_blk * BN + c0n + lane);
            __syncwarp();
          }
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(ra[i]);
          if (MODE == kBiasBf16 || MODE == kBiasGeluBf16) {
            if (has_bias) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(s_biasw + ew * 32 + i);
                const float2 lo = add2(make_float2(v[i], v[i + 1]), make_float2(b4.x, b4.y));
                const float2 hi = add2(make_float2(v[i + 2], v[i + 3]), make_float2(b4.z, b4.w));
                v[i] = lo.x; v[i + 1] = lo.y; v[i + 2] = hi.x; v[i + 3] = hi.y;
              }
            }
          }
          if (MODE == kGeluGradBf16 || MODE == kRowDotBf16) {
            const int ab = ci & 1;               // chunks 0 and 2 use buffer 0, chunk 1 buffer 1
            if (ci == 1 && slot + 2 * kEpiSlots < BN / Cfg::kColsPerChunk) {
              // third block of this tile: its aux goes into buffer 0 as soon as block 0's store has read it
              if (lane == 0) {
                tma_store_wait_read<0>();
                const int c0p = (slot + 2 * kEpiSlots) * Cfg::kColsPerChunk;
                mbar_arrive_expect_tx(&aux_bar[ew * 2], Cfg::kEpiBufBytes);
                tma_load_2d(my_epi, &tmC2, &aux_bar[ew * 2], n_blk * BN + c0p, row0);
              }
              __syncwarp();
            }
            mbar_wait(&aux_bar[ew * 2 + ab], (aux_phase >> ab) & 1u);
            aux_phase ^= (1u << ab);
            buf = ab;                            // the output block goes back into the buffer the aux came in
            const uint8_t* abuf = my_epi + ab * Cfg::kEpiBufBytes;
            // rows >= M arrive as zeros (TMA zero-fills out-of-bounds rows; their accumulator rows are zero as well)
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 u = *reinterpret_cast<const uint4*>(abuf + lane * 64 + ((uint32_t(i) ^ swz) << 4));
              const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                // bf16 -> fp32 is a 16-bit shift
                if (MODE == kGeluGradBf16) {
                  v[i * 8 + 2 * j] *= __uint_as_float(w[j] << 16);
                  v[i * 8 + 2 * j + 1] *= __uint_as_float(w[j] & 0xffff0000u);
                } else {
                  // D = rowsum(dO o O) over the bf16-rounded dO the attention backward will read
                  const float d0 = __bfloat162float(__float2bfloat16(v[i * 8 + 2 * j]));
                  const float d1 = __bfloat162float(__float2bfloat16(v[i * 8 + 2 * j + 1]));
                  dot = fmaf(d0, __uint_as_float(w[j] << 16), dot);
                  dot = fmaf(d1, __uint_as_float(w[j] & 0xffff0000u), dot);
                }
              }
            }
            if (MODE == kRowDotBf16) {
              // this 32-column block is half of head (col0 / 64): two atomics per (token, head) in total
              if (row < p.M) {
                const int clip = row / p.aux_period, tok = row - clip * p.aux_period;
                const int npad = ((p.aux_period + 127) >> 7) << 7;
                atomicAdd(const_cast<float*>(p.bias) + (size_t(clip) * (p.N >> 6) + (col0 >> 6)) * npad + tok, dot);
              }
            }
            if (MODE == kGeluGradBf16 && p.bias != nullptr) {
              // bias gradient of the layer that produced `pre`: column sums of this 32x32 block (rows >= M are 0).
              // One plain store per (quadrant, column): every such slot is written exactly once per tile.
              float cs[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) cs[i] = v[i];
              const float t = warp_colsum32(cs, lane);
              s_part[((tile_parity * 4 + q) << 8) + c0 + lane] = t;
            }
          }
          float w2[32];
          if (MODE == kBiasGeluBf16) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float2 g, dg;
              gelu_and_grad2(make_float2(v[i], v[i + 1]), g, dg);
              v[i] = dg.x; v[i + 1] = dg.y;
              w2[i] = g.x; w2[i + 1] = g.y;
            }
          }
          if (MODE != kGeluGradBf16 && MODE != kRowDotBf16) {
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
          }
          uint8_t* sbuf = my_epi + buf * Cfg::kEpiBufBytes;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint4 o;
            o.x = pack_bf16(v[ch * 8 + 0], v[ch * 8 + 1]);
            o.y = pack_bf16(v[ch * 8 + 2], v[ch * 8 + 3]);
            o.z = pack_bf16(v[ch * 8 + 4], v[ch * 8 + 5]);
            o.w = pack_bf16(v[ch * 8 + 6], v[ch * 8 + 7]);
            *reinterpret_cast<uint4*>(sbuf + lane * 64 + ((uint32_t(ch) ^ swz) << 4)) = o;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, sbuf, col0, row0);
            tma_store_commit();
          }
          buf ^= 1;
          if (MODE == kBiasGeluBf16) {
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
            uint8_t* sbuf2 = my_epi + buf * Cfg::kEpiBufBytes;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              uint4 o;
              o.x = pack_bf16(w2[ch * 8 + 0], w2[ch * 8 + 1]);
              o.y = pack_bf16(w2[ch * 8 + 2], w2[ch * 8 + 3]);
              o.z = pack_bf16(w2[ch * 8 + 4], w2[ch * 8 + 5]);
              o.w = pack_bf16(w2[ch * 8 + 6], w2[ch * 8 + 7]);
              *reinterpret_cast<uint4*>(sbuf2 + lane * 64 + ((uint32_t(ch) ^ swz) << 4)) = o;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC2, sbuf2, col0, row0);
              tma_store_commit();
            }
            buf ^= 1;
          }
        } else {
          uint32_t ra[16];
          tmem_ld_x16(taddr + c0, ra);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(ra[i]);
          if (MODE == kRowTabF32) {
            float4 add_cur[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) add_cur[i] = add_next[i];
            const int c0n = c0 + kEpiSlots * Cfg::kColsPerChunk;      // this warp's next chunk of the tile
            if (c0n < BN) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                if (add_row != nullptr) a = __ldg(reinterpret_cast<const float4*>(add_row + c0n) + i);
                if (p.bias != nullptr) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n_blk * BN + c0n) + i);
                  a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w;
                }
                add_next[i] = a;
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              v[4 * i] += add_cur[i].x; v[4 * i + 1] += add_cur[i].y;
              v[4 * i + 2] += add_cur[i].z; v[4 * i + 3] += add_cur[i].w;
            }
          }
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          uint8_t* sbuf = my_epi + buf * Cfg::kEpiBufBytes;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            float4 o = make_float4(v[ch * 4], v[ch * 4 + 1], v[ch * 4 + 2], v[ch * 4 + 3]);
            *reinterpret_cast<float4*>(sbuf + lane * 64 + ((uint32_t(ch) ^ swz) << 4)) = o;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            if (MODE == kWgradF32) tma_reduce_add_2d(&tmC, sbuf, col0, row0);
            else tma_store_2d(&tmC, sbuf, col0, row0);
            tma_store_commit();
          }
          buf ^= 1;
        }
      }
      if (MODE == kGeluGradBf16 && p.bias != nullptr) {
        // all 12 epilogue warps have stored their partial sums: add the four quadrants and flush this tile's 256
        // column sums.  The partials are double-buffered by tile parity, so the next tile's stores cannot overtake.
        named_bar_sync(2, kEpiWarps * 32);
        const int et = threadIdx.x - 64;
        if (et < BN) {
          const float* pp = &s_part[(tile_parity * 4) << 8];
          const float t = (pp[et] + pp[256 + et]) + (pp[512 + et] + pp[768 + et]);
          atomicAdd(const_cast<float*>(p.bias) + n_blk * BN + et, t);
        }
        tile_parity ^= 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();     // both CTAs finished using TMEM / each other's barriers
  if (warp == 1) tmem_dealloc_2cta<Cfg::kTmemCols>(tmem_base);
}

template <int MODE, bool BMN>
static int launch_gemm2_t(const void* A, const void* B, void* C, void* C2, const float* bias, const void* aux, int M,
                          int N, int K, int lda, int ldb, int ldc, int aux_period, int ld_aux, int splits,
                          cudaStream_t stream) {
  using Cfg = Gemm2Cfg<MODE, BMN>;
  constexpr int BN = Cfg::BN;
  if (N % BN != 0 || (lda % 8) || (ldb % 8)) return PB_ERR_BAD_ARG;
  CUtensorMap tmA, tmB, tmC, tmC2;
  int rc;
  if (!Cfg::kAMn) {
    if (K % 8) return PB_ERR_BAD_ARG;
    if ((rc = make_tmap_2d(&tmA, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, K, uint64_t(lda) * 2, BM, BK,
                           CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  } else {
    if (M % 256 != 0) return PB_ERR_BAD_ARG;
    if ((rc = make_tmap_2d(&tmA, A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, M, uint64_t(lda) * 2, BK, 64,
                           CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  }
  if (!Cfg::kBMn) {
    if ((rc = make_tmap_2d(&tmB, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, N, K, uint64_t(ldb) * 2, BN / 2, BK,
                           CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  } else {
    if ((rc = make_tmap_2d(&tmB, B, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, N, uint64_t(ldb) * 2, BK, 64,
                           CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  }
  if (Cfg::kOutF32) {
    if ((rc = make_tmap_2d(&tmC, C, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, M, N, uint64_t(ldc) * 4, 32, 16,
                           CU_TENSOR_MAP_SWIZZLE_64B)))
      return rc;
    tmC2 = tmC;
  } else {
    if ((rc = make_tmap_2d(&tmC, C, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, uint64_t(ldc) * 2, 32, 32,
                           CU_TENSOR_MAP_SWIZZLE_64B)))
      return rc;
    if (MODE == kBiasGeluBf16) {
      if ((rc = make_tmap_2d(&tmC2, C2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, uint64_t(ldc) * 2, 32, 32,
                             CU_TENSOR_MAP_SWIZZLE_64B)))
        return rc;
    } else if (MODE == kGeluGradBf16 || MODE == kRowDotBf16) {
      if (aux == nullptr || (ld_aux % 8)) return PB_ERR_BAD_ARG;
      if (MODE == kRowDotBf16 && (bias == nullptr || aux_period <= 0 || (N % 64) != 0)) return PB_ERR_BAD_ARG;
      if ((rc = make_tmap_2d(&tmC2, aux, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, M, N, uint64_t(ld_aux) * 2, 32, 32,
                             CU_TENSOR_MAP_SWIZZLE_64B)))
        return rc;
    } else {
      tmC2 = tmC;
    }
  }
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + 255) / 256;
  p.n_tiles = N / BN;
  p.k_blocks = (K + BK - 1) / BK;
  p.splits = Cfg::kWgrad ? (splits < 1 ? 1 : splits) : 1;
  if (p.splits > p.k_blocks) p.splits = p.k_blocks;
  {
    int per = (p.k_blocks + p.splits - 1) / p.splits;
    p.splits = (p.k_blocks + per - 1) / per;
  }
  p.bias = bias; p.aux = aux; p.aux_period = aux_period > 0 ? aux_period : 1; p.ld_aux = ld_aux;
  // K-major SW128: 8-row atoms of 1024 B, K advance of 16 elements = 32 B inside the swizzle atom.
  // MN-major SW128: 64-element MN groups 8192 B apart (LBO), 8-row K groups 1024 B apart (SBO), K advance = 2048 B.
  if (!Cfg::kAMn) { p.lbo_a = 16; p.sbo_a = 1024; p.kstep_a = 32; } else { p.lbo_a = 8192; p.sbo_a = 1024; p.kstep_a = 2048; }
  if (!Cfg::kBMn) { p.lbo_b = 16; p.sbo_b = 1024; p.kstep_b = 32; } else { p.lbo_b = 8192; p.sbo_b = 1024; p.kstep_b = 2048; }
  if (g_desc_override.active) {
    p.lbo_a = g_desc_override.v[0]; p.sbo_a = g_desc_override.v[1]; p.kstep_a = g_desc_override.v[2];
    p.lbo_b = g_desc_override.v[3]; p.sbo_b = g_desc_override.v[4]; p.kstep_b = g_desc_override.v[5];
  }
  PB_SET_SMEM_ONCE(Cfg::kSmemBytes, gemm2_kernel<MODE, BMN>);
  const int num_tiles = p.m_tiles * p.n_tiles * p.splits;
  int clusters = num_tiles < g_sm_limit / 2 ? num_tiles : g_sm_limit / 2;
  if (clusters <= 0) return 0;
  PB_LAUNCH((gemm2_kernel<MODE, BMN>), clusters * 2, kGemmThreads, Cfg::kSmemBytes, stream, tmA, tmB, tmC, tmC2, p);
  return 0;
}

int launch_gemm2(int mode, const void* A, const void* B, void* C, void* C2, const float* bias, const void* aux, int M,
                 int N, int K, int lda, int ldb, int ldc, int aux_period, int ld_aux, int splits, cudaStream_t stream) {
#define PB_G2(MODE_, BMN_) \
  launch_gemm2_t<MODE_, BMN_>(A, B, C, C2, bias, aux, M, N, K, lda, ldb, ldc, aux_period, ld_aux, splits, stream)
  switch (mode) {
    case kBiasBf16: return PB_G2(kBiasBf16, false);
    case kBiasGeluBf16: return PB_G2(kBiasGeluBf16, false);
    case kRowTabF32: return PB_G2(kRowTabF32, false);
    case kRowTabF32 | kBRowMajorKN: return PB_G2(kRowTabF32, true);   // fp32-tier dgrad: B is the row-stacked split weight
    case kGeluGradBf16: return PB_G2(kGeluGradBf16, false);
    case kWgradF32: return PB_G2(kWgradF32, false);
    case kBiasBf16 | kBRowMajorKN: return PB_G2(kBiasBf16, true);
    case kGeluGradBf16 | kBRowMajorKN: return PB_G2(kGeluGradBf16, true);
    case kRowDotBf16 | kBRowMajorKN: return PB_G2(kRowDotBf16, true);
    default: return PB_ERR_BAD_ARG;
  }
#undef PB_G2
}

}  // namespace pb
