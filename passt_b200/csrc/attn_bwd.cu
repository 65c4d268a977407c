// passt_b200 — fused multi-head attention backward on tcgen05 / TMEM (sm_100a), head_dim 64, non-causal.
//
// Autograd of Attention.forward's softmax(q k^T * scale) v (reference models/passt.py:345-358), recomputing the
// probabilities from the saved log-sum-exp instead of reading a stored [B,H,N,N] tensor.
//
// One CTA owns one (128-key tile, head, clip) and loops over the query tiles.  Everything is computed in the
// transposed (keys x queries) frame so that the tiles TMA brings in are used as-is by every MMA:
//     S^T  = K  Q^T        (A = K  K-major,            B = Q  K-major)        -> TMEM [0,128)
//     dP^T = V  dO^T       (A = V  K-major,            B = dO K-major)        -> TMEM [128,256)
//     P^T  = exp2(S^T c - lse[q]),  dS^T = P^T (dP^T - D[q])      (8 compute warps, thread = key row)
//     dV  += P^T  dO       (A = P^T  K-major (smem),   B = dO MN-major: same bytes as above)  -> TMEM [256,320)
//     dK  += dS^T Q        (A = dS^T K-major (smem),   B = Q  MN-major)                        -> TMEM [320,384)
//     dQ_i = dS   K        (A = dS^T read MN-major,    B = K  MN-major)                        -> TMEM [384,448)
// dQ_i tiles from different key tiles are summed in an fp32 buffer with TMA reduce-add; a small kernel then
// scales and packs dQ into the dqkv tensor.  D = rowsum(dO * O) comes from a pre-pass.
#include "common.cuh"

namespace pb {

constexpr int kBHd = 64;
constexpr int kBwdThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 compute
constexpr int kTile = 128;

struct AttnBwdParams {
  int N, H;
  float scale_log2, scale;
  const float* lse;   // [B,H,N]
  const float* Dsum;  // [B,H,N]
};

struct AttnBwdSmem {
  static constexpr int kK = 0;
  static constexpr int kV = kK + 16384;
  static constexpr int kQdO = kV + 16384;            // 2 stages x (Q 16 KB + dO 16 KB)
  static constexpr int kPT = kQdO + 2 * 32768;       // 32 KB
  static constexpr int kdST = kPT + 32768;           // 32 KB
  static constexpr int kdQ = kdST + 32768;           // 32 KB fp32 staging
  static constexpr int kVec = kdQ + 32768;           // lse/D: 2 x 2 x 128 floats
  static constexpr int kBars = kVec + 2048;
  static constexpr int kTotal = kBars + 128;
};

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmdO,
                const __grid_constant__ CUtensorMap tmdQKV, const __grid_constant__ CUtensorMap tmdQacc,
                const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sK = smem + AttnBwdSmem::kK;
  uint8_t* sV = smem + AttnBwdSmem::kV;
  uint8_t* sQdO = smem + AttnBwdSmem::kQdO;
  uint8_t* sPT = smem + AttnBwdSmem::kPT;
  uint8_t* sdST = smem + AttnBwdSmem::kdST;
  uint8_t* sdQ = smem + AttnBwdSmem::kdQ;
  float* sVec = reinterpret_cast<float*>(smem + AttnBwdSmem::kVec);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnBwdSmem::kBars);
  uint64_t* kv_full = bars;          // [1]
  uint64_t* qdo_full = bars + 1;     // [2]
  uint64_t* qdo_empty = bars + 3;    // [2]
  uint64_t* sdp_full = bars + 5;     // [1]
  uint64_t* pds_full = bars + 6;     // [1] 256 arrivals
  uint64_t* dq_full = bars + 7;      // [1]
  uint64_t* dkv_full = bars + 8;     // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * kTile;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_q = (p.N + kTile - 1) / kTile;
  const int C = p.H * kBHd;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmdQKV); tma_prefetch_desc(&tmdQacc);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 256);
    mbar_init(dq_full, 1);
    mbar_init(dkv_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320,
                 tdQ = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * 16384);
      tma_load_3d(sK, &tmQKV, kv_full, C + h * kBHd, kv0, b);
      tma_load_3d(sV, &tmQKV, kv_full, 2 * C + h * kBHd, kv0, b);
      for (int i = 0; i < n_q; ++i) {
        const int s = i & 1;
        mbar_wait(&qdo_empty[s], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&qdo_full[s], 2 * 16384);
        tma_load_3d(sQdO + s * 32768, &tmQKV, &qdo_full[s], h * kBHd, i * kTile, b);
        tma_load_3d(sQdO + s * 32768 + 16384, &tmdO, &qdo_full[s], h * kBHd, i * kTile, b);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t id_kv = make_idesc_bf16(128, 64, 0, 1);
    constexpr uint32_t id_q = make_idesc_bf16(128, 64, 1, 1);
    mbar_wait(kv_full, 0);
    for (int i = 0; i < n_q; ++i) {
      const int s = i & 1;
      mbar_wait(&qdo_full[s], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV);
      const uint32_t aQ = smem_u32(sQdO + s * 32768), adO = aQ + 16384;
      const uint32_t aPT = smem_u32(sPT), adST = smem_u32(sdST);
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tS, make_smem_desc_sw128(aK + k * 32, 16, 1024), make_smem_desc_sw128(aQ + k * 32, 16, 1024),
                       id_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tdP, make_smem_desc_sw128(aV + k * 32, 16, 1024),
                       make_smem_desc_sw128(adO + k * 32, 16, 1024), id_s, k > 0);
        tc_commit(sdp_full);
      }
      __syncwarp();
      mbar_wait(pds_full, i & 1);
      tc_fence_after();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t a_off = (k >> 2) * 16384 + (k & 3) * 32;   // K-major A, contraction = queries
          umma_bf16_ss(tdV, make_smem_desc_sw128(aPT + a_off, 16, 1024),
                       make_smem_desc_sw128(adO + k * 2048, 8192, 1024), id_kv, (i > 0 || k > 0));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t a_off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_bf16_ss(tdK, make_smem_desc_sw128(adST + a_off, 16, 1024),
                       make_smem_desc_sw128(aQ + k * 2048, 8192, 1024), id_kv, (i > 0 || k > 0));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)   // contraction = keys: dS^T rows; A is MN-major with two 64-query groups
          umma_bf16_ss(tdQ, make_smem_desc_sw128(adST + k * 2048, 16384, 1024),
                       make_smem_desc_sw128(aK + k * 2048, 8192, 1024), id_q, k > 0);
        tc_commit(dq_full);
        tc_commit(&qdo_empty[s]);
        if (i == n_q - 1) tc_commit(dkv_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== compute warps =====================
    const int cw = warp - 2;
    const int q = warp & 3;             // TMEM lane quadrant
    const int hc = cw >> 2;             // which half of the 128 columns this warp handles
    const int r = q * 32 + lane;        // row inside the tile (key row for S^T/dP^T, query row for dQ)
    const int ct = threadIdx.x - 64;    // 0..255
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const float log2e = 1.4426950408889634f;
    for (int i = 0; i < n_q; ++i) {
      const int q0 = i * kTile;
      float* s_lse = sVec + (i & 1) * 256;
      float* s_D = s_lse + 128;
      if (ct == 0) tma_store_wait_read<0>();   // dQ staging of the previous step has been read
      if (ct < 128) {
        const int qq = q0 + ct;
        s_lse[ct] = qq < p.N ? p.lse[(size_t(b) * p.H + h) * p.N + qq] * log2e : INFINITY;
      } else {
        const int qq = q0 + ct - 128;
        s_D[ct - 128] = qq < p.N ? p.Dsum[(size_t(b) * p.H + h) * p.N + qq] : 0.f;
      }
      named_bar_sync(1, 256);
      mbar_wait(sdp_full, i & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col0 = hc * 64 + c * 32;
        uint32_t sv[32], dv[32];
        tmem_ld_x32(tS + lane_addr + col0, sv);
        tmem_ld_x32(tdP + lane_addr + col0, dv);
        tmem_ld_wait();
        float pt[32], ds[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float pe = exp2f(__uint_as_float(sv[e]) * p.scale_log2 - s_lse[col0 + e]);
          pt[e] = pe;
          ds[e] = pe * (__uint_as_float(dv[e]) - s_D[col0 + e]);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int ch = c * 4 + g;   // 16-byte chunk inside this 64-query half
          const uint32_t off = hc * 16384 + r * 128 + ((ch ^ (r & 7)) << 4);
          uint4 o;
          o.x = pack_bf16(pt[g * 8 + 0], pt[g * 8 + 1]); o.y = pack_bf16(pt[g * 8 + 2], pt[g * 8 + 3]);
          o.z = pack_bf16(pt[g * 8 + 4], pt[g * 8 + 5]); o.w = pack_bf16(pt[g * 8 + 6], pt[g * 8 + 7]);
          *reinterpret_cast<uint4*>(sPT + off) = o;
          o.x = pack_bf16(ds[g * 8 + 0], ds[g * 8 + 1]); o.y = pack_bf16(ds[g * 8 + 2], ds[g * 8 + 3]);
          o.z = pack_bf16(ds[g * 8 + 4], ds[g * 8 + 5]); o.w = pack_bf16(ds[g * 8 + 6], ds[g * 8 + 7]);
          *reinterpret_cast<uint4*>(sdST + off) = o;
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(pds_full);
      // ---- drain dQ_i (rows = queries) -> fp32 staging -> TMA reduce-add into the accumulation buffer
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      {
        uint32_t v[32];
        tmem_ld_x32(tdQ + lane_addr + hc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          float4 o = make_float4(__uint_as_float(v[ch * 4]), __uint_as_float(v[ch * 4 + 1]),
                                 __uint_as_float(v[ch * 4 + 2]), __uint_as_float(v[ch * 4 + 3]));
          *reinterpret_cast<float4*>(sdQ + hc * 16384 + r * 128 + ((ch ^ (r & 7)) << 4)) = o;
        }
      }
      tc_fence_before();
      fence_proxy_async();
      named_bar_sync(1, 256);
      if (ct == 0) {
        tma_reduce_add_3d(&tmdQacc, sdQ, h * kBHd, q0, b);
        tma_reduce_add_3d(&tmdQacc, sdQ + 16384, h * kBHd + 32, q0, b);
        tma_store_commit();
      }
    }
    // ---- epilogue: dK (scaled), dV -> bf16 -> staging (P^T / dS^T buffers are free) -> TMA store
    mbar_wait(dkv_full, 0);
    tc_fence_after();
    {
      uint32_t vv[32], kk[32];
      tmem_ld_x32(tdV + lane_addr + hc * 32, vv);
      tmem_ld_x32(tdK + lane_addr + hc * 32, kk);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int ch = hc * 4 + g;
        const uint32_t off = r * 128 + ((ch ^ (r & 7)) << 4);
        uint4 o;
        o.x = pack_bf16(__uint_as_float(vv[g * 8 + 0]), __uint_as_float(vv[g * 8 + 1]));
        o.y = pack_bf16(__uint_as_float(vv[g * 8 + 2]), __uint_as_float(vv[g * 8 + 3]));
        o.z = pack_bf16(__uint_as_float(vv[g * 8 + 4]), __uint_as_float(vv[g * 8 + 5]));
        o.w = pack_bf16(__uint_as_float(vv[g * 8 + 6]), __uint_as_float(vv[g * 8 + 7]));
        *reinterpret_cast<uint4*>(sPT + off) = o;
        o.x = pack_bf16(__uint_as_float(kk[g * 8 + 0]) * p.scale, __uint_as_float(kk[g * 8 + 1]) * p.scale);
        o.y = pack_bf16(__uint_as_float(kk[g * 8 + 2]) * p.scale, __uint_as_float(kk[g * 8 + 3]) * p.scale);
        o.z = pack_bf16(__uint_as_float(kk[g * 8 + 4]) * p.scale, __uint_as_float(kk[g * 8 + 5]) * p.scale);
        o.w = pack_bf16(__uint_as_float(kk[g * 8 + 6]) * p.scale, __uint_as_float(kk[g * 8 + 7]) * p.scale);
        *reinterpret_cast<uint4*>(sdST + off) = o;
      }
    }
    tc_fence_before();
    fence_proxy_async();
    named_bar_sync(1, 256);
    if (ct == 0) {
      tma_store_3d(&tmdQKV, sdST, C + h * kBHd, kv0, b);       // dK
      tma_store_3d(&tmdQKV, sPT, 2 * C + h * kBHd, kv0, b);    // dV
      tma_store_commit();
      tma_store_wait<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// D[b,h,n] = sum_d dO[b,n,h,d] * O[b,n,h,d]     (one warp per token, 8 lanes per head)
__global__ void __launch_bounds__(256)
attn_dsum_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dO, float* __restrict__ Dsum,
                 int B, int N, int H) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= B * N) return;
  const int b = gw / N, n = gw - b * N;
  const int C = H * kBHd;
  const size_t off = size_t(gw) * C;
  for (int c0 = 0; c0 < C; c0 += 256) {
    const int col = c0 + lane * 8;
    float s = 0.f;
    if (col < C) {
      const uint4 a = *reinterpret_cast<const uint4*>(o + off + col);
      const uint4 g = *reinterpret_cast<const uint4*>(dO + off + col);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw4[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162*>(&aw[j]);
        const __nv_bfloat162 y = *reinterpret_cast<const __nv_bfloat162*>(&gw4[j]);
        s += __low2float(x) * __low2float(y) + __high2float(x) * __high2float(y);
      }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if ((lane & 7) == 0 && col < C) Dsum[(size_t(b) * H + col / kBHd) * N + n] = s;
  }
}

// dqkv[:, :, 0:C] = bf16(scale * dq_acc)
__global__ void __launch_bounds__(256)
attn_dq_pack_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, size_t rows, int C, float scale) {
  const size_t idx = (size_t(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (idx >= rows * C) return;
  const size_t row = idx / C;
  const int col = int(idx - row * C);
  const float4 a = *reinterpret_cast<const float4*>(acc + idx);
  const float4 c = *reinterpret_cast<const float4*>(acc + idx + 4);
  uint4 o;
  o.x = pack_bf16(a.x * scale, a.y * scale); o.y = pack_bf16(a.z * scale, a.w * scale);
  o.z = pack_bf16(c.x * scale, c.y * scale); o.w = pack_bf16(c.z * scale, c.w * scale);
  *reinterpret_cast<uint4*>(dqkv + row * size_t(3 * C) + col) = o;
}

}  // namespace pb

extern "C" {

size_t passt_attn_bwd_workspace_bytes(int B, int N, int H) {
  return size_t(B) * N * H * 64 * 4 + size_t(B) * H * N * 4 + 256;
}

// qkv bf16 [B,N,3C], o bf16 [B,N,C], dO bf16 [B,N,C], lse fp32 [B,H,N] -> dqkv bf16 [B,N,3C]
int passt_attn_bwd(const void* qkv, const void* o, const void* dO, const float* lse, void* dqkv, void* workspace,
                   int B, int N, int H, float scale, void* stream) {
  using namespace pb;
  if (B <= 0 || N <= 0 || H <= 0 || !workspace) return PB_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int C = H * kBHd;
  float* dq_acc = reinterpret_cast<float*>(workspace);
  float* Dsum = dq_acc + size_t(B) * N * C;
  PB_CUDA_TRY(cudaMemsetAsync(dq_acc, 0, size_t(B) * N * C * 4, st));
  {
    const long long warps = (long long)B * N;
    attn_dsum_kernel<<<int((warps * 32 + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)o,
                                                                    (const __nv_bfloat16*)dO, Dsum, B, N, H);
    PB_LAUNCH_CHECK();
  }
  CUtensorMap tmQKV, tmdO, tmdQKV, tmdQacc;
  int rc;
  if ((rc = make_tmap_3d(&tmQKV, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3 * C, N, B, uint64_t(3 * C) * 2,
                         uint64_t(N) * 3 * C * 2, kBHd, kTile, 1, CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  if ((rc = make_tmap_3d(&tmdO, dO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, C, N, B, uint64_t(C) * 2,
                         uint64_t(N) * C * 2, kBHd, kTile, 1, CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  if ((rc = make_tmap_3d(&tmdQKV, dqkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3 * C, N, B, uint64_t(3 * C) * 2,
                         uint64_t(N) * 3 * C * 2, kBHd, kTile, 1, CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  if ((rc = make_tmap_3d(&tmdQacc, dq_acc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, C, N, B, uint64_t(C) * 4,
                         uint64_t(N) * C * 4, 32, kTile, 1, CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  AttnBwdParams p;
  p.N = N; p.H = H; p.scale = scale; p.scale_log2 = scale * 1.4426950408889634f; p.lse = lse; p.Dsum = Dsum;
  static bool attr_set = false;
  if (!attr_set) {
    PB_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     AttnBwdSmem::kTotal));
    attr_set = true;
  }
  dim3 grid((N + kTile - 1) / kTile, H, B);
  attn_bwd_kernel<<<grid, kBwdThreads, AttnBwdSmem::kTotal, st>>>(tmQKV, tmdO, tmdQKV, tmdQacc, p);
  PB_LAUNCH_CHECK();
  {
    const size_t total = size_t(B) * N * C;
    attn_dq_pack_kernel<<<unsigned((total / 8 + 255) / 256), 256, 0, st>>>(dq_acc, (__nv_bfloat16*)dqkv,
                                                                          size_t(B) * N, C, scale);
    PB_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
