// passt_b200 — fused multi-head attention forward on tcgen05 / TMEM (sm_100a), head_dim 64, non-causal.
//
// Replaces Attention.forward's q@k^T * scale -> softmax -> @v (reference models/passt.py:345-358) without ever
// materialising the [B, H, N, N] score tensor.  Input is the packed qkv GEMM output [B, N, 3*H*64] (bf16) exactly
// as nn.Linear(dim, 3*dim) lays it out (:345 reshape(B, N, 3, H, hd)); output is [B, N, H*64] (bf16), i.e. the
// (attn @ v).transpose(1, 2).reshape(B, N, C) layout (:358), plus the per-row log-sum-exp for the backward pass.
//
// One CTA = one (128-query tile, head, clip).  Warp roles:
//   warp 0    : TMA producer (Q once, K/V tiles through a 2-stage ring; OOB rows are zero-filled by TMA)
//   warp 1    : TMEM allocator + tcgen05.mma issuer:  S = Q K^T (128x128x64),  O_j = P_j V_j (128x64x128)
//   warps 2-5 : softmax, one thread per query row (tcgen05.ld 32x32b): running max / sum in registers,
//               P written to shared memory in the 128B-swizzled K-major layout the PV MMA reads,
//               O accumulated in registers with the usual online-softmax rescale.
// TMEM: S at columns [0,128), O_j at [128,192) -> 256 columns allocated, two CTAs co-reside per SM so one CTA's
// softmax overlaps the other's MMAs.
#include "common.cuh"

namespace pb {

constexpr int kHd = 64;
constexpr int kAttnThreads = 192;
constexpr int kQTile = 128;
constexpr int kKvTile = 128;

struct AttnFwdParams {
  int N, H;
  float scale_log2;  // softmax scale * log2(e)
  float scale;
  float* lse;        // [B, H, N] natural-log LSE of the scaled scores
};

struct AttnFwdSmem {
  static constexpr int kQ = 0;
  static constexpr int kKV = kQ + kQTile * kHd * 2;                  // 2 stages x (K 16 KB + V 16 KB)
  static constexpr int kP = kKV + 2 * 2 * kKvTile * kHd * 2;         // P: 2 k-halves x 16 KB
  static constexpr int kBars = kP + 2 * kQTile * 64 * 2;
  static constexpr int kTotal = kBars + 128;
};

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                const AttnFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B operands need 1024-byte aligned tiles
  uint8_t* sQ = smem + AttnFwdSmem::kQ;
  uint8_t* sKV = smem + AttnFwdSmem::kKV;
  uint8_t* sP = smem + AttnFwdSmem::kP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnFwdSmem::kBars);
  uint64_t* q_full = bars;            // [1]
  uint64_t* kv_full = bars + 1;       // [2]
  uint64_t* kv_empty = bars + 3;      // [2]
  uint64_t* s_full = bars + 5;        // [1]
  uint64_t* p_full = bars + 6;        // [1]  (128 softmax threads arrive)
  uint64_t* o_full = bars + 7;        // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQTile;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.N + kKvTile - 1) / kKvTile;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;
  const int C = p.H * kHd;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kQTile * kHd * 2);
      tma_load_3d(sQ, &tmQKV, q_full, h * kHd, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        uint8_t* sK = sKV + s * (2 * kKvTile * kHd * 2);
        uint8_t* sV = sK + kKvTile * kHd * 2;
        mbar_arrive_expect_tx(&kv_full[s], 2 * kKvTile * kHd * 2);
        tma_load_3d(sK, &tmQKV, &kv_full[s], C + h * kHd, j * kKvTile, b);
        tma_load_3d(sV, &tmQKV, &kv_full[s], 2 * C + h * kHd, j * kKvTile, b);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // S = Q K^T : A,B K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);    // O = P V   : A K-major, B (V) MN-major
    auto issue_s = [&](int j) {
      const int s = j & 1;
      mbar_wait(&kv_full[s], (j >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t aQ = smem_u32(sQ);
        const uint32_t aK = smem_u32(sKV + s * (2 * kKvTile * kHd * 2));
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k)
          umma_bf16_ss(tmem_S, make_smem_desc_sw128(aQ + k * 32, 16, 1024), make_smem_desc_sw128(aK + k * 32, 16, 1024),
                       idesc_s, k > 0 ? 1u : 0u);
        tc_commit(s_full);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_s(0);
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      mbar_wait(p_full, j & 1);   // P_j in smem, S_j fully consumed, O_{j-1} drained
      tc_fence_after();
      // S_{j+1} first so the softmax warps of this CTA can start on it while PV_j runs
      if (j + 1 < n_kv) issue_s(j + 1);
      if (lane == 0) {
        const uint32_t aP = smem_u32(sP);
        const uint32_t aV = smem_u32(sKV + s * (2 * kKvTile * kHd * 2) + kKvTile * kHd * 2);
#pragma unroll
        for (int k = 0; k < kKvTile / 16; ++k) {
          const uint64_t da = make_smem_desc_sw128(aP + (k >> 2) * (kQTile * 128) + (k & 3) * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(aV + k * 2048, 8192, 1024);
          umma_bf16_ss(tmem_O, da, db, idesc_o, k > 0 ? 1u : 0u);
        }
        tc_commit(o_full);
        tc_commit(&kv_empty[s]);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax / accumulate warps =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;            // query row inside the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    float o_acc[kHd];
#pragma unroll
    for (int i = 0; i < kHd; ++i) o_acc[i] = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int kv_valid = min(kKvTile, p.N - j * kKvTile);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tmem_S + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f((m_run - m_new) * p.scale_log2);
      const float moff = m_new * p.scale_log2;
      // drain O_{j-1} (issued one iteration ago) before P is overwritten / O is recomputed
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t v[16];
          tmem_ld_x16(tmem_O + lane_addr + c * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o_acc[c * 16 + i] += __uint_as_float(v[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < kHd; ++i) o_acc[i] *= alpha;
      // pass 2: P = exp2(S*c - m*c) -> bf16 -> swizzled smem
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tmem_S + lane_addr + c * 32, v);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = exp2f(__uint_as_float(v[i]) * p.scale_log2 - moff);
          pv[i] = (c * 32 + i < kv_valid) ? e : 0.f;
          rs += pv[i];
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16(pv[g * 8 + 0], pv[g * 8 + 1]);
          o.y = pack_bf16(pv[g * 8 + 2], pv[g * 8 + 3]);
          o.z = pack_bf16(pv[g * 8 + 4], pv[g * 8 + 5]);
          o.w = pack_bf16(pv[g * 8 + 6], pv[g * 8 + 7]);
          const int col = c * 32 + g * 8;          // key index of this 16-byte chunk
          const int kh = col >> 6, ch = (col & 63) >> 3;
          *reinterpret_cast<uint4*>(sP + kh * (kQTile * 128) + r * 128 + ((ch ^ (r & 7)) << 4)) = o;
        }
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(p_full);
    }
    // last partial O
    mbar_wait(o_full, (n_kv - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t v[16];
      tmem_ld_x16(tmem_O + lane_addr + c * 16, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) o_acc[c * 16 + i] += __uint_as_float(v[i]);
    }
    const float inv_l = 1.0f / l_run;
    if (q0 + r < p.N)
      p.lse[(size_t(b) * p.H + h) * p.N + q0 + r] = m_run * p.scale + logf(l_run);
    // O tile -> swizzled staging (the P buffer is free now: PV_last has completed) -> TMA store
    uint8_t* stage = sP + q * 4096;
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
    if (lane == 0 && q0 + q * 32 < p.N) {
      tma_store_3d(&tmO, stage, h * kHd, q0 + q * 32, b);
      tma_store_commit();
      tma_store_wait<0>();
    }
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
  p.N = N; p.H = H; p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f; p.lse = lse;
  static bool attr_set = false;
  if (!attr_set) {
    PB_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnFwdSmem::kTotal));
    attr_set = true;
  }
  dim3 grid((N + kQTile - 1) / kQTile, H, B);
  attn_fwd_kernel<<<grid, kAttnThreads, AttnFwdSmem::kTotal, reinterpret_cast<cudaStream_t>(stream)>>>(tmQKV, tmO, p);
  PB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
