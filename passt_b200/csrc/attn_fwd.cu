// passt_b200 — fused multi-head attention forward on tcgen05 / TMEM (sm_100a), head_dim 64, non-causal.
//
// Replaces Attention.forward's q@k^T * scale -> softmax -> @v (reference models/passt.py:345-358) without ever
// materialising the [B, H, N, N] score tensor.  Input is the packed qkv GEMM output [B, N, 3*H*64] (bf16) exactly
// as nn.Linear(dim, 3*dim) lays it out (:345 reshape(B, N, 3, H, hd)); output is [B, N, H*64] (bf16), i.e. the
// (attn @ v).transpose(1, 2).reshape(B, N, C) layout (:358), plus the per-row log-sum-exp for the backward pass.
//
// One CTA = one (128-query tile, head, clip).  Warp roles:
//   warp 0    : TMA producer (Q once, K/V tiles through a 2-stage ring; OOB rows are zero-filled by TMA)
//   warp 1    : TMEM allocator + tcgen05.mma issuer
//   warps 2-5 : softmax, one thread per query row (tcgen05.ld 32x32b)
// Both MMAs take their A operand from TENSOR MEMORY (the M=128 x N<=128 shapes are operand-bandwidth bound when A
// comes from shared memory):
//   S   = Q K^T  : A = Q copied once into TMEM (32 columns of packed bf16), B = K tile (smem, K-major)
//   O_j = P_j V_j: A = P_j written by the softmax threads with tcgen05.st over the first 64 columns of the S
//                  accumulator they have just consumed, B = V tile (smem, MN-major view of the [keys, hd] tile)
// Online softmax (running max / sum) and the O accumulator live in registers.
// TMEM columns: S/P [0,128)  O_j [128,192)  Q [192,224)  -> 256 allocated, two CTAs co-reside per SM so one CTA's
// softmax overlaps the other's MMAs.
#include "common.cuh"

namespace pb {

constexpr int kHd = 64;
constexpr int kAttnThreads = 192;
constexpr int kQTile = 128;
constexpr int kKvTile = 128;

struct AttnFwdParams {
  int N, H, B;
  int n_qt;          // query tiles per (clip, head)
  int total_items;   // B * H * n_qt
  float scale_log2;  // softmax scale * log2(e)
  float scale;
  float* lse;        // [B, H, N] natural-log LSE of the scaled scores
};

struct AttnFwdSmem {
  static constexpr int kQ = 0;                                       // 2 Q tiles (ring); reused as output staging
  static constexpr int kKV = kQ + 2 * kQTile * kHd * 2;              // 2 stages x (K 16 KB + V 16 KB)
  static constexpr int kBars = kKV + 2 * 2 * kKvTile * kHd * 2;
  static constexpr int kTotal = kBars + 128;
};

// Persistent: each CTA walks a strided list of (query tile, head, clip) items.  The TMA producer runs ahead across
// item boundaries (next Q tile + next K/V tiles), so the load latency and the per-CTA prologue are paid once.
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                const AttnFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B operands need 1024-byte aligned tiles
  uint8_t* sQ = smem + AttnFwdSmem::kQ;
  uint8_t* sKV = smem + AttnFwdSmem::kKV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnFwdSmem::kBars);
  uint64_t* q_full = bars;            // [2]  TMA -> softmax warps
  uint64_t* q_empty = bars + 2;       // [2]  4 arrivals (one per softmax warp, after its output store was read)
  uint64_t* kv_full = bars + 4;       // [2]
  uint64_t* kv_empty = bars + 6;      // [2]
  uint64_t* s_full = bars + 8;        // [1]
  uint64_t* p_full = bars + 9;        // [1]  (128 softmax threads arrive)
  uint64_t* o_full = bars + 10;       // [1]
  uint64_t* q_ready = bars + 11;      // [1]  (128 arrivals: Q is in TMEM)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (p.N + kKvTile - 1) / kKvTile;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 4);
      mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(q_ready, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tmem_S = tmem_base;         // fp32 scores, 128 columns
  const uint32_t tmem_P = tmem_base;         // bf16 probabilities over the consumed scores, 64 columns
  const uint32_t tmem_O = tmem_base + 128;   // fp32 partial output, 64 columns
  const uint32_t tmem_Q = tmem_base + 192;   // bf16 queries, 32 columns
  const int C = p.H * kHd;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t t = 0, n = 0;
      for (int it = blockIdx.x; it < p.total_items; it += gridDim.x, ++n) {
        const int qt = it % p.n_qt, h = (it / p.n_qt) % p.H, b = it / (p.n_qt * p.H);
        const uint32_t qs = n & 1;
        mbar_wait(&q_empty[qs], ((n >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qs], kQTile * kHd * 2);
        tma_load_3d(sQ + qs * (kQTile * kHd * 2), &tmQKV, &q_full[qs], h * kHd, qt * kQTile, b);
        for (int j = 0; j < n_kv; ++j, ++t) {
          const uint32_t s = t & 1;
          mbar_wait(&kv_empty[s], ((t >> 1) & 1) ^ 1);
          uint8_t* sK = sKV + s * (2 * kKvTile * kHd * 2);
          uint8_t* sV = sK + kKvTile * kHd * 2;
          mbar_arrive_expect_tx(&kv_full[s], 2 * kKvTile * kHd * 2);
          tma_load_3d(sK, &tmQKV, &kv_full[s], C + h * kHd, j * kKvTile, b);
          tma_load_3d(sV, &tmQKV, &kv_full[s], 2 * C + h * kHd, j * kKvTile, b);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // S = Q K^T : A TMEM, B K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);    // O = P V   : A TMEM, B (V) MN-major
    auto issue_s = [&](uint32_t t) {
      const uint32_t s = t & 1;
      mbar_wait(&kv_full[s], (t >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t aK = smem_u32(sKV + s * (2 * kKvTile * kHd * 2));
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k)
          umma_bf16_ts(tmem_S, tmem_Q + k * 8, make_smem_desc_sw128(aK + k * 32, 16, 1024), idesc_s,
                       k > 0 ? 1u : 0u);
        tc_commit(s_full);
      }
      __syncwarp();
    };
    uint32_t t = 0, n = 0;
    for (int it = blockIdx.x; it < p.total_items; it += gridDim.x, ++n) {
      mbar_wait(q_ready, n & 1);
      tc_fence_after();
      issue_s(t);
      for (int j = 0; j < n_kv; ++j, ++t) {
        const uint32_t s = t & 1;
        mbar_wait(p_full, t & 1);   // P_t is in TMEM, the previous O has been drained
        tc_fence_after();
        if (lane == 0) {
          const uint32_t aV = smem_u32(sKV + s * (2 * kKvTile * kHd * 2) + kKvTile * kHd * 2);
#pragma unroll
          for (int k = 0; k < kKvTile / 16; ++k)
            umma_bf16_ts(tmem_O, tmem_P + k * 8, make_smem_desc_sw128(aV + k * 2048, 8192, 1024), idesc_o,
                         k > 0 ? 1u : 0u);
          tc_commit(o_full);
          tc_commit(&kv_empty[s]);
        }
        __syncwarp();
        // S_{t+1} overwrites the columns P_t lives in: issue order (after PV_t) keeps it safe
        if (j + 1 < n_kv) issue_s(t + 1);
      }
    }
  } else {
    // ===================== softmax / accumulate warps =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;            // query row inside the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    // Q tile of item n: swizzled smem row -> packed bf16 in TMEM (A operand of S = Q K^T)
    auto stage_q = [&](uint32_t n) {
      const uint32_t qs = n & 1;
      mbar_wait(&q_full[qs], (n >> 1) & 1);
      const uint8_t* src = sQ + qs * (kQTile * kHd * 2);
      uint32_t qv[32];
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 u = *reinterpret_cast<const uint4*>(src + r * 128 + ((ch ^ (r & 7)) << 4));
        qv[ch * 4] = u.x; qv[ch * 4 + 1] = u.y; qv[ch * 4 + 2] = u.z; qv[ch * 4 + 3] = u.w;
      }
      tmem_st_x32(tmem_Q + lane_addr, qv);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(q_ready);
    };
    uint32_t t = 0, n = 0;
    if (blockIdx.x < p.total_items) stage_q(0);
    for (int it = blockIdx.x; it < p.total_items; it += gridDim.x, ++n) {
      const int qt = it % p.n_qt, h = (it / p.n_qt) % p.H, b = it / (p.n_qt * p.H);
      const int q0 = qt * kQTile;
      float m_run = -INFINITY, l_run = 0.f;
      float o_acc[kHd];
#pragma unroll
      for (int i = 0; i < kHd; ++i) o_acc[i] = 0.f;

      for (int j = 0; j < n_kv; ++j, ++t) {
        const int kv_valid = min(kKvTile, p.N - j * kKvTile);
        const bool full_tile = (kv_valid == kKvTile);     // warp-uniform: only the last key tile needs masking
        mbar_wait(s_full, t & 1);
        tc_fence_after();
        // pass 1: row max
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_x32(tmem_S + lane_addr + c * 32, v);
          tmem_ld_wait();
          if (full_tile) {
#pragma unroll
            for (int i = 0; i < 32; i += 2)
              mx = fmaxf(mx, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
        const float m_new = fmaxf(m_run, mx);
        const float alpha = ex2_approx((m_run - m_new) * p.scale_log2);
        const float moff = m_new * p.scale_log2;
        // drain the previous partial O: S_t was issued after PV_{t-1}, so it is complete (the wait is immediate)
        if (j > 0) {
          mbar_wait(o_full, (t - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_x32(tmem_O + lane_addr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = (o_acc[c * 32 + i] + __uint_as_float(v[i])) * alpha;
          }
        }
        // pass 2: P = exp2(S*c - m*c) -> packed bf16 -> TMEM, over the score columns already consumed
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_x32(tmem_S + lane_addr + c * 32, v);
          tmem_ld_wait();
          uint32_t pk[16];
          if (full_tile) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float e0 = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -moff));
              const float e1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -moff));
              rs0 += e0;
              rs1 += e1;
              pk[i >> 1] = pack_bf16(e0, e1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float e0 = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2, -moff));
              float e1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), p.scale_log2, -moff));
              e0 = (c * 32 + i < kv_valid) ? e0 : 0.f;
              e1 = (c * 32 + i + 1 < kv_valid) ? e1 : 0.f;
              rs0 += e0;
              rs1 += e1;
              pk[i >> 1] = pack_bf16(e0, e1);
            }
          }
          tmem_st_x16(tmem_P + lane_addr + c * 16, pk);
        }
        l_run = l_run * alpha + (rs0 + rs1);
        m_run = m_new;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full);
      }
      // every S MMA of this item has completed (s_full of its last tile): the Q columns are free -> stage the next Q
      // now so the MMA warp can start the next item while this one's epilogue runs
      const bool has_next = (it + int(gridDim.x) < p.total_items);
      if (has_next) stage_q(n + 1);
      // last partial O
      mbar_wait(o_full, (t - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tmem_O + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] += __uint_as_float(v[i]);
      }
      const float inv_l = 1.0f / l_run;
      if (q0 + r < p.N)
        p.lse[(size_t(b) * p.H + h) * p.N + q0 + r] = m_run * p.scale + logf(l_run);
      // O tile -> swizzled staging (this item's Q buffer: Q has lived in TMEM since the item started) -> TMA store
      uint8_t* stage = sQ + (n & 1) * (kQTile * kHd * 2) + q * 4096;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint4 o;
        o.x = pack_bf16(o_acc[ch * 8 + 0] * inv_l, o_acc[ch * 8 + 1] * inv_l);
        o.y = pack_bf16(o_acc[ch * 8 + 2] * inv_l, o_acc[ch * 8 + 3] * inv_l);
        o.z = pack_bf16(o_acc[ch * 8 + 4] * inv_l, o_acc[ch * 8 + 5] * inv_l);
        o.w = pack_bf16(o_acc[ch * 8 + 6] * inv_l, o_acc[ch * 8 + 7] * inv_l);
        *reinterpret_cast<uint4*>(stage + lane * 128 + ((ch ^ (lane & 7)) << 4)) = o;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (q0 + q * 32 < p.N) {
          tma_store_3d(&tmO, stage, h * kHd, q0 + q * 32, b);
          tma_store_commit();
          tma_store_wait_read<0>();     // the Q buffer may be refilled once the store has read it
        }
        mbar_arrive(&q_empty[n & 1]);
      }
      __syncwarp();
    }
    if (lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

}  // namespace pb

extern "C" {

// qkv: bf16 [B, N, 3*H*64]; out: bf16 [B, N, H*64]; lse: fp32 [B, H, N]
int passt_attn_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, float scale, void* stream) {
  using namespace pb;
  if (B <= 0 || N <= 0 || H <= 0) return PB_ERR_BAD_ARG;
  const int C = H * kHd;
  CUtensorMap tmQKV, tmO;
  int rc;
  if ((rc = make_tmap_3d(&tmQKV, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3 * C, N, B, uint64_t(3 * C) * 2,
                         uint64_t(N) * 3 * C * 2, kHd, kKvTile, 1, CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  if ((rc = make_tmap_3d(&tmO, out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, C, N, B, uint64_t(C) * 2,
                         uint64_t(N) * C * 2, kHd, 32, 1, CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  AttnFwdParams p;
  p.N = N; p.H = H; p.B = B; p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f; p.lse = lse;
  p.n_qt = (N + kQTile - 1) / kQTile;
  p.total_items = B * H * p.n_qt;
  static bool attr_set = false;
  if (!attr_set) {
    PB_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnFwdSmem::kTotal));
    attr_set = true;
  }
  const int grid = p.total_items < 2 * kNumSMs ? p.total_items : 2 * kNumSMs;
  attn_fwd_kernel<<<grid, kAttnThreads, AttnFwdSmem::kTotal, reinterpret_cast<cudaStream_t>(stream)>>>(tmQKV, tmO, p);
  PB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
