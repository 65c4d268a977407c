// passt_b200 — fused multi-head attention backward on tcgen05 / TMEM (sm_100a), head_dim 64, non-causal.
//
// Autograd of Attention.forward's softmax(q k^T * scale) v (reference models/passt.py:345-358), recomputing the
// probabilities from the saved log-sum-exp instead of reading a stored [B,H,N,N] tensor.
//
// One CTA owns one (128-key tile, head, clip) and loops over the query tiles.  Everything is computed in the
// transposed (keys x queries) frame so that the tiles TMA brings in are used as-is by every MMA:
//     S^T  = K  Q^T        (A = K  in TMEM,            B = Q  K-major)        -> TMEM [0,128)
//     dP^T = V  dO^T       (A = V  in TMEM,            B = dO K-major)        -> TMEM [128,256)
//     P^T  = exp2(S^T c - lse[q]),  dS^T = P^T (dP^T - D[q])      (8 compute warps, thread = key row)
//     dV  += P^T  dO       (A = P^T  in TMEM,          B = dO MN-major: same smem bytes as above) -> TMEM [256,320)
//     dK  += dS^T Q        (A = dS^T in TMEM,          B = Q  MN-major)                           -> TMEM [320,384)
//     dQ_i = dS   K        (A = dS^T smem MN-major,    B = K  MN-major)                           -> TMEM [384,448)
// M=128 x N<=128 MMAs are operand-bandwidth bound when A is read from shared memory, so every A operand that can
// live in tensor memory does: K and V are copied there once per CTA (columns [448,512)), P^T / dS^T are written by
// the compute warps with tcgen05.st over the S^T / dP^T columns they have just consumed (bf16 pairs, half the
// width).  Only dQ keeps a shared-memory A operand (it needs dS with queries on the M axis).
// dQ_i tiles from different key tiles are summed in an fp32 buffer with TMA reduce-add; a small kernel then
// scales and packs dQ into the dqkv tensor.  D = rowsum(dO * O) comes from a pre-pass.
#include "common.cuh"

namespace pb {

constexpr int kBHd = 64;
constexpr int kBwdThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 compute
constexpr int kTile = 128;

struct AttnBwdParams {
  int N, H;
  int n_kvt;         // key tiles per (clip, head)
  int total_items;   // B * H * n_kvt
  float scale_log2, scale;
  const float* lse;   // [B,H,Npad] log2 domain, pad rows +inf
  const float* Dsum;  // [B,H,Npad] pad rows 0
  float* dq_acc;      // [B,N,C] fp32 accumulation of dQ over key tiles
  float* dbias;       // [3C] qkv bias gradient (+=) or nullptr: column sums of dK / dV are folded into the epilogue
  int Npad;
  long long* timeline;  // bring-up only: clock64 stamps of CTA 0 (nullptr in production)
};
#define PB_STAMP(role, idx)                                                                     \
  do {                                                                                           \
    if (p.timeline != nullptr && blockIdx.x == 0 && (idx) < 512) p.timeline[(role) * 512 + (idx)] = clock64(); \
  } while (0)

struct AttnBwdSmem {
  static constexpr int kKV = 0;                      // 2 item buffers x (K 16 KB + V 16 KB)
  static constexpr int kQdO = kKV + 2 * 32768;       // 2 stages x (Q 16 KB + dO 16 KB)
  static constexpr int kdST = kQdO + 2 * 32768;      // 32 KB  dS^T (bf16) for the dQ MMA
  static constexpr int kdQ = kdST + 32768;           // 32 KB  fp32 dQ staging; dK/dV staging at item end
  static constexpr int kVec = kdQ + 32768;           // lse/D: 2 x 2 x 128 floats
  static constexpr int kBias = kVec + 2048;          // per-CTA partial qkv-bias gradient (K and V thirds): 2 x 768 floats
  static constexpr int kBars = kBias + 2 * 768 * 4;
  static constexpr int kTotal = kBars + 256;
};

// 32 lanes x 32 values -> lane i ends up with the sum over all lanes of value i (butterfly: 31 shuffles)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < n / 2) {
        const float send = up ? v[j] : v[j + n / 2];
        const float recv = __shfl_xor_sync(0xffffffffu, send, off);
        v[j] = (up ? v[j + n / 2] : v[j]) + recv;
      }
    }
  }
  return v[0];
}

// Persistent: each CTA walks a strided list of (key tile, head, clip) items; the TMA producer prefetches the next
// item's K/V and Q/dO tiles while the current item is still being processed.
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmdO,
                const __grid_constant__ CUtensorMap tmdQKV, const __grid_constant__ CUtensorMap tmdQacc,
                const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sKV = smem + AttnBwdSmem::kKV;
  uint8_t* sQdO = smem + AttnBwdSmem::kQdO;
  uint8_t* sdST = smem + AttnBwdSmem::kdST;
  uint8_t* sdQ = smem + AttnBwdSmem::kdQ;
  float* sVec = reinterpret_cast<float*>(smem + AttnBwdSmem::kVec);
  float* sBias = reinterpret_cast<float*>(smem + AttnBwdSmem::kBias);   // [0,768): dK sums, [768,1536): dV sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnBwdSmem::kBars);
  uint64_t* kv_full = bars;          // [2]
  uint64_t* kv_empty = bars + 2;     // [2]  tcgen05.commit after the item's last MMA
  uint64_t* qdo_full = bars + 4;     // [2]
  uint64_t* qdo_empty = bars + 6;    // [2]
  uint64_t* sdp_full = bars + 8;     // [1]
  uint64_t* pds_full = bars + 9;     // [1] 256 arrivals
  uint64_t* dq_full = bars + 10;     // [1]
  uint64_t* dkv_full = bars + 11;    // [1]
  uint64_t* kv_ready = bars + 12;    // [1] 256 arrivals: K, V copied into TMEM
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_q = (p.N + kTile - 1) / kTile;
  const int C = p.H * kBHd;
  // the last query tile is trimmed to whole 32-query chunks: S^T / dP^T run with N = qcols_last, the math skips the
  // unused score columns and dV / dK contract over qcols_last queries (N = 474: 96 instead of 128)
  const int qcols_last = ((p.N - (n_q - 1) * kTile + 31) / 32) * 32;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmdQKV); tma_prefetch_desc(&tmdQacc);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1);
      mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 256);
    mbar_init(dq_full, 1);
    mbar_init(dkv_full, 1);
    mbar_init(kv_ready, 256);
    fence_barrier_init();
  }
  if (p.dbias != nullptr)
    for (int i = threadIdx.x; i < 2 * 768; i += blockDim.x) sBias[i] = 0.f;
  if (warp == 1) tmem_alloc<512>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_gate();
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320,
                 tdQ = tmem_base + 384, tK = tmem_base + 448, tV = tmem_base + 480;
  // P^T (bf16 pairs) for query half hc lives at tS + 64*hc .. +32, dS^T at tdP + 64*hc .. +32: each compute warp
  // only overwrites score columns it has itself already read

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0, n = 0;
      for (int it = blockIdx.x; it < p.total_items; it += gridDim.x, ++n) {
        const int kvt = it % p.n_kvt, h = (it / p.n_kvt) % p.H, b = it / (p.n_kvt * p.H);
        const uint32_t kb = n & 1;
        mbar_wait(&kv_empty[kb], ((n >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[kb], 2 * 16384);
        tma_load_3d(sKV + kb * 32768, &tmQKV, &kv_full[kb], C + h * kBHd, kvt * kTile, b);
        tma_load_3d(sKV + kb * 32768 + 16384, &tmQKV, &kv_full[kb], 2 * C + h * kBHd, kvt * kTile, b);
        for (int i = 0; i < n_q; ++i, ++g) {
          const uint32_t s = g & 1;
          mbar_wait(&qdo_empty[s], ((g >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&qdo_full[s], 2 * 16384 + 2 * 512);
          tma_load_3d(sQdO + s * 32768, &tmQKV, &qdo_full[s], h * kBHd, i * kTile, b);
          tma_load_3d(sQdO + s * 32768 + 16384, &tmdO, &qdo_full[s], h * kBHd, i * kTile, b);
          const size_t voff = (size_t(b) * p.H + h) * p.Npad + size_t(i) * kTile;
          bulk_load_1d(sVec + s * 256, p.lse + voff, 512, &qdo_full[s]);
          bulk_load_1d(sVec + s * 256 + 128, p.Dsum + voff, 512, &qdo_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t id_kv = make_idesc_bf16(128, 64, 0, 1);
    constexpr uint32_t id_q = make_idesc_bf16(128, 64, 1, 1);
    // base descriptors are computed once, warp-uniformly; inside the loop an elected lane only adds the per-k offset
    // (the start-address field is in 16-byte units) and issues the tcgen05 instructions
    uint64_t bQk[2], bdOk[2], bQmn[2], bdOmn[2], bKmn[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const uint32_t aQ = smem_u32(sQdO + s * 32768), adO = aQ + 16384;
      bQk[s] = make_smem_desc_sw128(aQ, 16, 1024);          // Q  as K-major B  (S^T = K Q^T)
      bdOk[s] = make_smem_desc_sw128(adO, 16, 1024);        // dO as K-major B  (dP^T = V dO^T)
      bQmn[s] = make_smem_desc_sw128(aQ, 8192, 1024);       // Q  as MN-major B (dK += dS^T Q)
      bdOmn[s] = make_smem_desc_sw128(adO, 8192, 1024);     // dO as MN-major B (dV += P^T dO)
      bKmn[s] = make_smem_desc_sw128(smem_u32(sKV + s * 32768), 8192, 1024);   // K as MN-major B (dQ = dS K)
    }
    const uint64_t bdST = make_smem_desc_sw128(smem_u32(sdST), 16384, 1024);   // dS^T as MN-major A (dQ = dS K)
    const uint32_t id_s_last = make_idesc_bf16(128, qcols_last, 0, 0);
    uint32_t g = 0, n = 0;
    for (int it = blockIdx.x; it < p.total_items; it += gridDim.x, ++n) {
      const uint32_t kb = n & 1;
      mbar_wait(kv_ready, n & 1);
      tc_fence_after();
      for (int i = 0; i < n_q; ++i, ++g) {
        const uint32_t s = g & 1;
        mbar_wait(&qdo_full[s], (g >> 1) & 1);
        tc_fence_after();
        if (lane == 0) PB_STAMP(1, g * 4 + 0);
        const bool last_q = (i == n_q - 1);
        const uint32_t ids = last_q ? id_s_last : id_s;
        const int n_qk = last_q ? qcols_last / 16 : 8;      // 16-query contraction steps of dV / dK
        if (elect_one()) {
          const uint64_t qk = s ? bQk[1] : bQk[0], dok = s ? bdOk[1] : bdOk[0];
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(tS, tK + k * 8, qk + uint64_t(k * 2), ids, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(tdP, tV + k * 8, dok + uint64_t(k * 2), ids, k > 0);
          tc_commit(sdp_full);
        }
        __syncwarp();
        if (lane == 0) PB_STAMP(1, g * 4 + 1);
        mbar_wait(pds_full, g & 1);
        tc_fence_after();
        if (lane == 0) PB_STAMP(1, g * 4 + 2);
        if (elect_one()) {
          const uint64_t qmn = s ? bQmn[1] : bQmn[0], domn = s ? bdOmn[1] : bdOmn[0];
          const uint64_t kmn = kb ? bKmn[1] : bKmn[0];
#pragma unroll
          for (int k = 0; k < 8; ++k)   // contraction = keys: dS^T rows; A is MN-major with two 64-query groups
            umma_bf16_ss(tdQ, bdST + uint64_t(k * 128), kmn + uint64_t(k * 128), id_q, k > 0);
          tc_commit(dq_full);           // dQ first: the compute warps drain it while dV / dK accumulate
#pragma unroll
          for (int k = 0; k < 8; ++k)   // contraction = queries; 16 queries = 8 TMEM columns of packed bf16
            if (k < n_qk)
              umma_bf16_ts(tdV, tS + (k >> 2) * 64 + (k & 3) * 8, domn + uint64_t(k * 128), id_kv, (i > 0 || k > 0));
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < n_qk)
              umma_bf16_ts(tdK, tdP + (k >> 2) * 64 + (k & 3) * 8, qmn + uint64_t(k * 128), id_kv, (i > 0 || k > 0));
          tc_commit(&qdo_empty[s]);
          if (i == n_q - 1) {
            tc_commit(dkv_full);
            tc_commit(&kv_empty[kb]);
          }
        }
        __syncwarp();
        if (lane == 0) PB_STAMP(1, g * 4 + 3);
      }
    }
  } else {
    // ===================== compute warps =====================
    const int cw = warp - 2;
    const int q = warp & 3;             // TMEM lane quadrant
    const int hc = cw >> 2;             // which half of the 128 columns this warp handles
    const int r = q * 32 + lane;        // row inside the tile (key row for S^T/dP^T, query row for dQ)
    const int ct = threadIdx.x - 64;    // 0..255
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    // K (warps with hc == 0) and V (hc == 1) rows of item n: swizzled smem -> packed bf16 in TMEM
    auto stage_kv = [&](uint32_t n) {
      const uint32_t kb = n & 1;
      mbar_wait(&kv_full[kb], (n >> 1) & 1);
      const uint8_t* src = sKV + kb * 32768 + (hc == 0 ? 0 : 16384);
      uint32_t kv[32];
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 u = *reinterpret_cast<const uint4*>(src + r * 128 + ((ch ^ (r & 7)) << 4));
        kv[ch * 4] = u.x; kv[ch * 4 + 1] = u.y; kv[ch * 4 + 2] = u.z; kv[ch * 4 + 3] = u.w;
      }
      tmem_st_x32((hc == 0 ? tK : tV) + lane_addr, kv);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(kv_ready);
    };
    uint32_t g = 0, n = 0;
    if (blockIdx.x < p.total_items) stage_kv(0);
    for (int it = blockIdx.x; it < p.total_items; it += gridDim.x, ++n) {
      const int kvt = it % p.n_kvt, h = (it / p.n_kvt) % p.H, b = it / (p.n_kvt * p.H);
      const int kv0 = kvt * kTile;
      const bool has_next = (it + int(gridDim.x) < p.total_items);
      for (int i = 0; i < n_q; ++i, ++g) {
        const int q0 = i * kTile;
        const float* s_lse = sVec + (g & 1) * 256;     // log2-domain LSE of this query tile (TMA bulk copy)
        const float* s_D = s_lse + 128;
        mbar_wait(&qdo_full[g & 1], (g >> 1) & 1);     // lse / D landed with the Q / dO tiles
        if (ct == 0) PB_STAMP(0, g * 6 + 0);
        mbar_wait(sdp_full, g & 1);
        tc_fence_after();
        if (ct == 0) PB_STAMP(0, g * 6 + 1);
        const int q_cols = (i == n_q - 1) ? qcols_last : kTile;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col0 = hc * 64 + c * 32;
          if (col0 >= q_cols) break;           // trimmed last query tile (warp-uniform)
          uint32_t sv[32], dv[32];
          tmem_ld_x32(tS + lane_addr + col0, sv);
          tmem_ld_x32(tdP + lane_addr + col0, dv);
          tmem_ld_wait();
          uint32_t pp[16], dd2[16];
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const float4 l4 = *reinterpret_cast<const float4*>(s_lse + col0 + e);
            const float4 d4 = *reinterpret_cast<const float4*>(s_D + col0 + e);
            const float p0 = ex2_approx(fmaf(__uint_as_float(sv[e]), p.scale_log2, -l4.x));
            const float p1 = ex2_approx(fmaf(__uint_as_float(sv[e + 1]), p.scale_log2, -l4.y));
            const float p2 = ex2_approx(fmaf(__uint_as_float(sv[e + 2]), p.scale_log2, -l4.z));
            const float p3 = ex2_approx(fmaf(__uint_as_float(sv[e + 3]), p.scale_log2, -l4.w));
            pp[e >> 1] = pack_bf16(p0, p1);
            pp[(e >> 1) + 1] = pack_bf16(p2, p3);
            dd2[e >> 1] = pack_bf16(p0 * (__uint_as_float(dv[e]) - d4.x), p1 * (__uint_as_float(dv[e + 1]) - d4.y));
            dd2[(e >> 1) + 1] =
                pack_bf16(p2 * (__uint_as_float(dv[e + 2]) - d4.z), p3 * (__uint_as_float(dv[e + 3]) - d4.w));
          }
          tmem_st_x16(tS + lane_addr + hc * 64 + c * 16, pp);     // A operand of dV += P^T dO
          tmem_st_x16(tdP + lane_addr + hc * 64 + c * 16, dd2);   // A operand of dK += dS^T Q
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {   // dS^T also goes to smem: dQ = dS K reads it MN-major
            const int ch = c * 4 + gq;       // 16-byte chunk inside this 64-query half
            const uint32_t off = hc * 16384 + r * 128 + ((ch ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(sdST + off) =
                make_uint4(dd2[gq * 4], dd2[gq * 4 + 1], dd2[gq * 4 + 2], dd2[gq * 4 + 3]);
          }
        }
        if (ct == 0) PB_STAMP(0, g * 6 + 2);
        tmem_st_wait();
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(pds_full);
        if (ct == 0) PB_STAMP(0, g * 6 + 3);
        // all S^T / dP^T MMAs of this item are complete after its last sdp_full: K/V columns are free -> stage the
        // next item's K/V now so its first MMAs overlap this item's tail
        if (i == n_q - 1 && has_next) stage_kv(n + 1);
        // ---- drain dQ_i (rows = queries) -> fp32 staging -> TMA reduce-add into the accumulation buffer
        mbar_wait(dq_full, g & 1);
        tc_fence_after();
        if (ct == 0) PB_STAMP(0, g * 6 + 4);
        if (ct == 0) tma_store_wait_read<0>();   // staging (dQ of the previous step / dK,dV of the previous item) was read
        named_bar_sync(1, 256);
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
          // TMA reduce-add of the dQ tile (rows = queries) into the fp32 accumulation buffer; rows >= N are clipped
          tma_reduce_add_3d(&tmdQacc, sdQ, h * kBHd, q0, b);
          tma_reduce_add_3d(&tmdQacc, sdQ + 16384, h * kBHd + 32, q0, b);
          tma_store_commit();
        }
        if (ct == 0) PB_STAMP(0, g * 6 + 5);
      }
      // ---- item epilogue: dK (scaled), dV -> bf16 -> staging (dQ staging buffer) -> TMA store
      mbar_wait(dkv_full, n & 1);
      tc_fence_after();
      if (ct == 0) tma_store_wait_read<0>();
      named_bar_sync(1, 256);
      {
        uint32_t vv[32], kk[32];
        tmem_ld_x32(tdV + lane_addr + hc * 32, vv);
        tmem_ld_x32(tdK + lane_addr + hc * 32, kk);
        tmem_ld_wait();
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const int ch = hc * 4 + gq;
          const uint32_t off = r * 128 + ((ch ^ (r & 7)) << 4);
          uint4 o;
          o.x = pack_bf16(__uint_as_float(vv[gq * 8 + 0]), __uint_as_float(vv[gq * 8 + 1]));
          o.y = pack_bf16(__uint_as_float(vv[gq * 8 + 2]), __uint_as_float(vv[gq * 8 + 3]));
          o.z = pack_bf16(__uint_as_float(vv[gq * 8 + 4]), __uint_as_float(vv[gq * 8 + 5]));
          o.w = pack_bf16(__uint_as_float(vv[gq * 8 + 6]), __uint_as_float(vv[gq * 8 + 7]));
          *reinterpret_cast<uint4*>(sdQ + off) = o;                                  // dV tile
          o.x = pack_bf16(__uint_as_float(kk[gq * 8 + 0]) * p.scale, __uint_as_float(kk[gq * 8 + 1]) * p.scale);
          o.y = pack_bf16(__uint_as_float(kk[gq * 8 + 2]) * p.scale, __uint_as_float(kk[gq * 8 + 3]) * p.scale);
          o.z = pack_bf16(__uint_as_float(kk[gq * 8 + 4]) * p.scale, __uint_as_float(kk[gq * 8 + 5]) * p.scale);
          o.w = pack_bf16(__uint_as_float(kk[gq * 8 + 6]) * p.scale, __uint_as_float(kk[gq * 8 + 7]) * p.scale);
          *reinterpret_cast<uint4*>(sdQ + 16384 + off) = o;                          // dK tile
        }
        if (p.dbias != nullptr) {
          // qkv bias gradient, K and V thirds: column sums of the bf16-rounded dK / dV rows of this tile
          const bool valid = (kv0 + r < p.N);
          float cv[32], ck[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            cv[e] = valid ? __bfloat162float(__float2bfloat16(__uint_as_float(vv[e]))) : 0.f;
            ck[e] = valid ? __bfloat162float(__float2bfloat16(__uint_as_float(kk[e]) * p.scale)) : 0.f;
          }
          const float sv = warp_colsum32(cv, lane);
          const float sk = warp_colsum32(ck, lane);
          atomicAdd(&sBias[768 + h * kBHd + hc * 32 + lane], sv);    // shared-memory partials, flushed once per CTA
          atomicAdd(&sBias[h * kBHd + hc * 32 + lane], sk);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      named_bar_sync(1, 256);
      if (ct == 0) {
        tma_store_3d(&tmdQKV, sdQ + 16384, C + h * kBHd, kv0, b);   // dK
        tma_store_3d(&tmdQKV, sdQ, 2 * C + h * kBHd, kv0, b);       // dV
        tma_store_commit();
      }
    }
    if (ct == 0) tma_store_wait<0>();
    if (p.dbias != nullptr) {
      named_bar_sync(1, 256);
      for (int i = ct; i < 2 * 768; i += 256) {
        const float v = sBias[i];
        if (v != 0.f) atomicAdd(p.dbias + C + i, v);     // K third at [C, 2C), V third at [2C, 3C)
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// Variant 2 (PASST_B200_ATTN_BWD=2): the same schedule with SIXTEEN compute warps -- four per TMEM lane quadrant, each
// owning a 32-query column quarter of the score tiles instead of a 64-query half.  The math phase between the S^T/dP^T
// MMAs and the dQ/dV/dK MMAs is a serial stretch of the step (tensor pipe idle); with two warps per scheduler its
// tcgen05.ld / MUFU / shared-memory latencies are exposed, with four they overlap.  Layout differences: the packed P^T
// / dS^T operand of quarter qc starts at its own first score column (TMEM column qc*32, 16 columns), so a warp only
// overwrites score columns it has read itself; K / V staging, the dQ drain and the dK / dV epilogue are split in 16-
// column pieces.
// ---------------------------------------------------------------------------------------------
constexpr int kBwd2Threads = 576;   // warp 0 TMA, warp 1 MMA, warps 2..17 compute

// 32 lanes x 16 values -> lanes l and l + 16 end up with the sum over all lanes of value (l & 15)
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int off = 8, n = 16; off >= 1; off >>= 1, n >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < n / 2) {
        const float send = up ? v[j] : v[j + n / 2];
        const float recv = __shfl_xor_sync(0xffffffffu, send, off);
        v[j] = (up ? v[j + n / 2] : v[j]) + recv;
      }
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}

__global__ void __launch_bounds__(kBwd2Threads, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmdO,
                 const __grid_constant__ CUtensorMap tmdQKV, const __grid_constant__ CUtensorMap tmdQacc,
                 const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sKV = smem + AttnBwdSmem::kKV;
  uint8_t* sQdO = smem + AttnBwdSmem::kQdO;
  uint8_t* sdST = smem + AttnBwdSmem::kdST;
  uint8_t* sdQ = smem + AttnBwdSmem::kdQ;
  float* sVec = reinterpret_cast<float*>(smem + AttnBwdSmem::kVec);
  float* sBias = reinterpret_cast<float*>(smem + AttnBwdSmem::kBias);   // [0,768): dK sums, [768,1536): dV sums
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnBwdSmem::kBars);
  uint64_t* kv_full = bars;          // [2]
  uint64_t* kv_empty = bars + 2;     // [2]  tcgen05.commit after the item's last MMA
  uint64_t* qdo_full = bars + 4;     // [2]
  uint64_t* qdo_empty = bars + 6;    // [2]
  uint64_t* sdp_full = bars + 8;     // [1]
  uint64_t* pds_full = bars + 9;     // [1] 512 arrivals
  uint64_t* dq_full = bars + 10;     // [1]
  uint64_t* dkv_full = bars + 11;    // [1]
  uint64_t* kv_ready = bars + 12;    // [1] 512 arrivals: K, V copied into TMEM
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 13);
  constexpr int kCT = 512;           // compute threads

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_q = (p.N + kTile - 1) / kTile;
  const int C = p.H * kBHd;
  const int qcols_last = ((p.N - (n_q - 1) * kTile + 31) / 32) * 32;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmdQKV); tma_prefetch_desc(&tmdQacc);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1);
      mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, kCT);
    mbar_init(dq_full, 1);
    mbar_init(dkv_full, 1);
    mbar_init(kv_ready, kCT);
    fence_barrier_init();
  }
  if (p.dbias != nullptr)
    for (int i = threadIdx.x; i < 2 * 768; i += blockDim.x) sBias[i] = 0.f;
  if (warp == 1) tmem_alloc<512>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_gate();
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320,
                 tdQ = tmem_base + 384, tK = tmem_base + 448, tV = tmem_base + 480;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0, n = 0;
      for (int it = blockIdx.x; it < p.total_items; it += gridDim.x, ++n) {
        const int kvt = it % p.n_kvt, h = (it / p.n_kvt) % p.H, b = it / (p.n_kvt * p.H);
        const uint32_t kb = n & 1;
        mbar_wait(&kv_empty[kb], ((n >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[kb], 2 * 16384);
        tma_load_3d(sKV + kb * 32768, &tmQKV, &kv_full[kb], C + h * kBHd, kvt * kTile, b);
        tma_load_3d(sKV + kb * 32768 + 16384, &tmQKV, &kv_full[kb], 2 * C + h * kBHd, kvt * kTile, b);
        for (int i = 0; i < n_q; ++i, ++g) {
          const uint32_t s = g & 1;
          mbar_wait(&qdo_empty[s], ((g >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&qdo_full[s], 2 * 16384 + 2 * 512);
          tma_load_3d(sQdO + s * 32768, &tmQKV, &qdo_full[s], h * kBHd, i * kTile, b);
          tma_load_3d(sQdO + s * 32768 + 16384, &tmdO, &qdo_full[s], h * kBHd, i * kTile, b);
          const size_t voff = (size_t(b) * p.H + h) * p.Npad + size_t(i) * kTile;
          bulk_load_1d(sVec + s * 256, p.lse + voff, 512, &qdo_full[s]);
          bulk_load_1d(sVec + s * 256 + 128, p.Dsum + voff, 512, &qdo_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t id_kv = make_idesc_bf16(128, 64, 0, 1);
    constexpr uint32_t id_q = make_idesc_bf16(128, 64, 1, 1);
    uint64_t bQk[2], bdOk[2], bQmn[2], bdOmn[2], bKmn[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const uint32_t aQ = smem_u32(sQdO + s * 32768), adO = aQ + 16384;
      bQk[s] = make_smem_desc_sw128(aQ, 16, 1024);          // Q  as K-major B  (S^T = K Q^T)
      bdOk[s] = make_smem_desc_sw128(adO, 16, 1024);        // dO as K-major B  (dP^T = V dO^T)
      bQmn[s] = make_smem_desc_sw128(aQ, 8192, 1024);       // Q  as MN-major B (dK += dS^T Q)
      bdOmn[s] = make_smem_desc_sw128(adO, 8192, 1024);     // dO as MN-major B (dV += P^T dO)
      bKmn[s] = make_smem_desc_sw128(smem_u32(sKV + s * 32768), 8192, 1024);   // K as MN-major B (dQ = dS K)
    }
    const uint64_t bdST = make_smem_desc_sw128(smem_u32(sdST), 16384, 1024);   // dS^T as MN-major A (dQ = dS K)
    const uint32_t id_s_last = make_idesc_bf16(128, qcols_last, 0, 0);
    uint32_t g = 0, n = 0;
    for (int it = blockIdx.x; it < p.total_items; it += gridDim.x, ++n) {
      const uint32_t kb = n & 1;
      mbar_wait(kv_ready, n & 1);
      tc_fence_after();
      for (int i = 0; i < n_q; ++i, ++g) {
        const uint32_t s = g & 1;
        mbar_wait(&qdo_full[s], (g >> 1) & 1);
        tc_fence_after();
        const bool last_q = (i == n_q - 1);
        const uint32_t ids = last_q ? id_s_last : id_s;
        const int n_qk = last_q ? qcols_last / 16 : 8;      // 16-query contraction steps of dV / dK
        if (elect_one()) {
          const uint64_t qk = s ? bQk[1] : bQk[0], dok = s ? bdOk[1] : bdOk[0];
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(tS, tK + k * 8, qk + uint64_t(k * 2), ids, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(tdP, tV + k * 8, dok + uint64_t(k * 2), ids, k > 0);
          tc_commit(sdp_full);
        }
        __syncwarp();
        mbar_wait(pds_full, g & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t qmn = s ? bQmn[1] : bQmn[0], domn = s ? bdOmn[1] : bdOmn[0];
          const uint64_t kmn = kb ? bKmn[1] : bKmn[0];
#pragma unroll
          for (int k = 0; k < 8; ++k)   // contraction = keys: dS^T rows; A is MN-major with two 64-query groups
            umma_bf16_ss(tdQ, bdST + uint64_t(k * 128), kmn + uint64_t(k * 128), id_q, k > 0);
          tc_commit(dq_full);           // dQ first: the compute warps drain it while dV / dK accumulate
#pragma unroll
          for (int k = 0; k < 8; ++k)   // contraction = queries; 16 queries = 8 TMEM columns of packed bf16,
            if (k < n_qk)               // quarter k/2 starts at column (k/2)*32
              umma_bf16_ts(tdV, tS + (k >> 1) * 32 + (k & 1) * 8, domn + uint64_t(k * 128), id_kv, (i > 0 || k > 0));
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < n_qk)
              umma_bf16_ts(tdK, tdP + (k >> 1) * 32 + (k & 1) * 8, qmn + uint64_t(k * 128), id_kv, (i > 0 || k > 0));
          tc_commit(&qdo_empty[s]);
          if (i == n_q - 1) {
            tc_commit(dkv_full);
            tc_commit(&kv_empty[kb]);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== compute warps =====================
    const int cw = warp - 2;            // 0..15
    const int q = warp & 3;             // TMEM lane quadrant
    const int qc = cw >> 2;             // which 32-column quarter of the 128 query columns this warp handles
    const int hc = qc >> 1;             // 64-query half the quarter belongs to (shared-memory operand groups)
    const int r = q * 32 + lane;        // row inside the tile (key row for S^T/dP^T, query row for dQ)
    const int ct = threadIdx.x - 64;    // 0..511
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    // K (quarters 0, 1) and V (quarters 2, 3) rows of item n, 32 of the 64 dims each: swizzled smem -> packed bf16 in TMEM
    auto stage_kv = [&](uint32_t n) {
      const uint32_t kb = n & 1;
      mbar_wait(&kv_full[kb], (n >> 1) & 1);
      const uint8_t* src = sKV + kb * 32768 + (qc < 2 ? 0 : 16384);
      const int half = qc & 1;
      uint32_t kv[16];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int cc = half * 4 + ch;
        const uint4 u = *reinterpret_cast<const uint4*>(src + r * 128 + ((cc ^ (r & 7)) << 4));
        kv[ch * 4] = u.x; kv[ch * 4 + 1] = u.y; kv[ch * 4 + 2] = u.z; kv[ch * 4 + 3] = u.w;
      }
      tmem_st_x16((qc < 2 ? tK : tV) + half * 16 + lane_addr, kv);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(kv_ready);
    };
    uint32_t g = 0, n = 0;
    if (blockIdx.x < p.total_items) stage_kv(0);
    for (int it = blockIdx.x; it < p.total_items; it += gridDim.x, ++n) {
      const int kvt = it % p.n_kvt, h = (it / p.n_kvt) % p.H, b = it / (p.n_kvt * p.H);
      const int kv0 = kvt * kTile;
      const bool has_next = (it + int(gridDim.x) < p.total_items);
      for (int i = 0; i < n_q; ++i, ++g) {
        const int q0 = i * kTile;
        const float* s_lse = sVec + (g & 1) * 256;     // log2-domain LSE of this query tile (TMA bulk copy)
        const float* s_D = s_lse + 128;
        mbar_wait(&qdo_full[g & 1], (g >> 1) & 1);     // lse / D landed with the Q / dO tiles
        mbar_wait(sdp_full, g & 1);
        tc_fence_after();
        const int q_cols = (i == n_q - 1) ? qcols_last : kTile;
        const int col0 = qc * 32;
        if (col0 < q_cols) {                 // trimmed last query tile: whole quarters drop out (warp-uniform)
          uint32_t pp[16], dd2[16];
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            uint32_t sv[16], dv[16];
            tmem_ld_x16(tS + lane_addr + col0 + sub * 16, sv);
            tmem_ld_x16(tdP + lane_addr + col0 + sub * 16, dv);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const float4 l4 = *reinterpret_cast<const float4*>(s_lse + col0 + sub * 16 + e);
              const float4 d4 = *reinterpret_cast<const float4*>(s_D + col0 + sub * 16 + e);
              const float p0 = ex2_approx(fmaf(__uint_as_float(sv[e]), p.scale_log2, -l4.x));
              const float p1 = ex2_approx(fmaf(__uint_as_float(sv[e + 1]), p.scale_log2, -l4.y));
              const float p2 = ex2_approx(fmaf(__uint_as_float(sv[e + 2]), p.scale_log2, -l4.z));
              const float p3 = ex2_approx(fmaf(__uint_as_float(sv[e + 3]), p.scale_log2, -l4.w));
              const int o = sub * 8 + (e >> 1);
              pp[o] = pack_bf16(p0, p1);
              pp[o + 1] = pack_bf16(p2, p3);
              dd2[o] = pack_bf16(p0 * (__uint_as_float(dv[e]) - d4.x), p1 * (__uint_as_float(dv[e + 1]) - d4.y));
              dd2[o + 1] = pack_bf16(p2 * (__uint_as_float(dv[e + 2]) - d4.z), p3 * (__uint_as_float(dv[e + 3]) - d4.w));
            }
          }
          // both 16-column loads of this quarter are complete: its first 16 columns take the packed operands
          tmem_st_x16(tS + lane_addr + col0, pp);      // A operand of dV += P^T dO
          tmem_st_x16(tdP + lane_addr + col0, dd2);    // A operand of dK += dS^T Q
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {   // dS^T also goes to smem: dQ = dS K reads it MN-major
            const int ch = (qc & 1) * 4 + gq;   // 16-byte chunk inside this 64-query half
            const uint32_t off = hc * 16384 + r * 128 + ((ch ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(sdST + off) =
                make_uint4(dd2[gq * 4], dd2[gq * 4 + 1], dd2[gq * 4 + 2], dd2[gq * 4 + 3]);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(pds_full);
        if (i == n_q - 1 && has_next) stage_kv(n + 1);
        // ---- drain dQ_i (rows = queries) -> fp32 staging -> TMA reduce-add into the accumulation buffer
        mbar_wait(dq_full, g & 1);
        tc_fence_after();
        if (ct == 0) tma_store_wait_read<0>();   // staging (dQ of the previous step / dK,dV of the previous item) was read
        named_bar_sync(1, kCT);
        {
          uint32_t v[16];
          tmem_ld_x16(tdQ + lane_addr + qc * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = (qc & 1) * 4 + j;
            float4 o = make_float4(__uint_as_float(v[j * 4]), __uint_as_float(v[j * 4 + 1]),
                                   __uint_as_float(v[j * 4 + 2]), __uint_as_float(v[j * 4 + 3]));
            *reinterpret_cast<float4*>(sdQ + hc * 16384 + r * 128 + ((ch ^ (r & 7)) << 4)) = o;
          }
        }
        tc_fence_before();
        fence_proxy_async();
        named_bar_sync(1, kCT);
        if (ct == 0) {
          tma_reduce_add_3d(&tmdQacc, sdQ, h * kBHd, q0, b);
          tma_reduce_add_3d(&tmdQacc, sdQ + 16384, h * kBHd + 32, q0, b);
          tma_store_commit();
        }
      }
      // ---- item epilogue: dK (scaled), dV -> bf16 -> staging (dQ staging buffer) -> TMA store
      mbar_wait(dkv_full, n & 1);
      tc_fence_after();
      if (ct == 0) tma_store_wait_read<0>();
      named_bar_sync(1, kCT);
      {
        uint32_t vv[16], kk[16];
        tmem_ld_x16(tdV + lane_addr + qc * 16, vv);
        tmem_ld_x16(tdK + lane_addr + qc * 16, kk);
        tmem_ld_wait();
#pragma unroll
        for (int gq = 0; gq < 2; ++gq) {
          const int ch = qc * 2 + gq;
          const uint32_t off = r * 128 + ((ch ^ (r & 7)) << 4);
          uint4 o;
          o.x = pack_bf16(__uint_as_float(vv[gq * 8 + 0]), __uint_as_float(vv[gq * 8 + 1]));
          o.y = pack_bf16(__uint_as_float(vv[gq * 8 + 2]), __uint_as_float(vv[gq * 8 + 3]));
          o.z = pack_bf16(__uint_as_float(vv[gq * 8 + 4]), __uint_as_float(vv[gq * 8 + 5]));
          o.w = pack_bf16(__uint_as_float(vv[gq * 8 + 6]), __uint_as_float(vv[gq * 8 + 7]));
          *reinterpret_cast<uint4*>(sdQ + off) = o;                                  // dV tile
          o.x = pack_bf16(__uint_as_float(kk[gq * 8 + 0]) * p.scale, __uint_as_float(kk[gq * 8 + 1]) * p.scale);
          o.y = pack_bf16(__uint_as_float(kk[gq * 8 + 2]) * p.scale, __uint_as_float(kk[gq * 8 + 3]) * p.scale);
          o.z = pack_bf16(__uint_as_float(kk[gq * 8 + 4]) * p.scale, __uint_as_float(kk[gq * 8 + 5]) * p.scale);
          o.w = pack_bf16(__uint_as_float(kk[gq * 8 + 6]) * p.scale, __uint_as_float(kk[gq * 8 + 7]) * p.scale);
          *reinterpret_cast<uint4*>(sdQ + 16384 + off) = o;                          // dK tile
        }
        if (p.dbias != nullptr) {
          const bool valid = (kv0 + r < p.N);
          float cv[16], ck[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            cv[e] = valid ? __bfloat162float(__float2bfloat16(__uint_as_float(vv[e]))) : 0.f;
            ck[e] = valid ? __bfloat162float(__float2bfloat16(__uint_as_float(kk[e]) * p.scale)) : 0.f;
          }
          const float sv = warp_colsum16(cv, lane);
          const float sk = warp_colsum16(ck, lane);
          if (lane < 16) {
            atomicAdd(&sBias[768 + h * kBHd + qc * 16 + lane], sv);    // shared-memory partials, flushed once per CTA
            atomicAdd(&sBias[h * kBHd + qc * 16 + lane], sk);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async();
      named_bar_sync(1, kCT);
      if (ct == 0) {
        tma_store_3d(&tmdQKV, sdQ + 16384, C + h * kBHd, kv0, b);   // dK
        tma_store_3d(&tmdQKV, sdQ, 2 * C + h * kBHd, kv0, b);       // dV
        tma_store_commit();
      }
    }
    if (ct == 0) tma_store_wait<0>();
    if (p.dbias != nullptr) {
      named_bar_sync(1, kCT);
      for (int i = ct; i < 2 * 768; i += kCT) {
        const float v = sBias[i];
        if (v != 0.f) atomicAdd(p.dbias + C + i, v);     // K third at [C, 2C), V third at [2C, 3C)
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

long long* g_attn_bwd_timeline = nullptr;
int g_attn_bwd_variant = -1;      // 1: eight compute warps (default), 2: sixteen; -1: read PASST_B200_ATTN_BWD on first use


// D[b,h,n] = sum_d dO[b,n,h,d] * O[b,n,h,d]     (one warp per token, 8 lanes per head)
__global__ void __launch_bounds__(256)
attn_dsum_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dO, float* __restrict__ Dsum,
                 int B, int N, int H, int Npad) {
  pdl_gate();
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
    if ((lane & 7) == 0 && col < C) Dsum[(size_t(b) * H + col / kBHd) * Npad + n] = s;
  }
}

// dqkv[:, :, 0:C] = bf16(scale * dq_acc); optionally dbias[0:C] += column sums of the packed values.
// grid (C/256, row chunks); block 256 = 32 column groups (8 columns each) x 8 rows in flight
__global__ void __launch_bounds__(256)
attn_dq_pack_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dbias,
                    int rows, int C, float scale, int rows_per_cta) {
  __shared__ float red[8][256];
  pdl_gate();
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + cg * 8;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float cs[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) cs[e] = 0.f;
  for (int r = r0 + rl; r < r1; r += 8) {
    const float* src = acc + size_t(r) * C + col;
    const float4 a = *reinterpret_cast<const float4*>(src);
    const float4 c = *reinterpret_cast<const float4*>(src + 4);
    uint4 o;
    o.x = pack_bf16(a.x * scale, a.y * scale); o.y = pack_bf16(a.z * scale, a.w * scale);
    o.z = pack_bf16(c.x * scale, c.y * scale); o.w = pack_bf16(c.z * scale, c.w * scale);
    *reinterpret_cast<uint4*>(dqkv + size_t(r) * (3 * C) + col) = o;
    const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
      cs[2 * j] += __low2float(h2);
      cs[2 * j + 1] += __high2float(h2);
    }
  }
  if (dbias != nullptr) {
#pragma unroll
    for (int e = 0; e < 8; ++e) red[rl][cg * 8 + e] = cs[e];
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(dbias + blockIdx.x * 256 + threadIdx.x, t);
  }
}

}  // namespace pb

extern "C" {

void passt_attn_bwd_debug_timeline(void* buf) { pb::g_attn_bwd_timeline = reinterpret_cast<long long*>(buf); }
/* 1: eight compute warps, 2: sixteen compute warps (32-query column quarters) */
void passt_attn_bwd_set_variant(int v) { pb::g_attn_bwd_variant = (v == 2) ? 2 : 1; }
int passt_attn_bwd_get_variant(void) { return pb::g_attn_bwd_variant; }

size_t passt_attn_bwd_workspace_bytes(int B, int N, int H) {
  const size_t npad = size_t((N + 127) / 128) * 128;
  return size_t(B) * N * H * 64 * 4 + size_t(B) * H * npad * 4 + 256;
}

// qkv bf16 [B,N,3C], o bf16 [B,N,C], dO bf16 [B,N,C], lse fp32 [B,H,Npad] (log2 domain, from passt_attn_fwd)
// -> dqkv bf16 [B,N,3C]; dbias_qkv (optional, fp32 [3C]) += column sums of dqkv (the qkv Linear's bias gradient)
// D-fusion support: zero the workspace (dQ accumulator + padded D) ahead of time and hand out the D buffer, so that the
// GEMM that produces dO (mode kRowDotBf16) can accumulate D = rowsum(dO o O) in its epilogue; then call
// passt_attn_bwd_ex with flags = 1 (workspace prepared, D already accumulated: no memset, no D pre-pass).
int passt_attn_bwd_prepare(void* workspace, int B, int N, int H, void* stream) {
  using namespace pb;
  if (B <= 0 || N <= 0 || H <= 0 || !workspace) return PB_ERR_BAD_ARG;
  const int C = H * kBHd;
  const int Npad = ((N + kTile - 1) / kTile) * kTile;
  PB_CUDA_TRY(cudaMemsetAsync(workspace, 0, size_t(B) * N * C * 4 + size_t(B) * H * Npad * 4,
                              reinterpret_cast<cudaStream_t>(stream)));
  return 0;
}
float* passt_attn_bwd_dsum_ptr(void* workspace, int B, int N, int H) {
  return reinterpret_cast<float*>(workspace) + size_t(B) * N * H * pb::kBHd;
}

int passt_attn_bwd_ex(const void* qkv, const void* o, const void* dO, const float* lse, void* dqkv, float* dbias_qkv,
                      void* workspace, int B, int N, int H, float scale, int flags, void* stream);

int passt_attn_bwd(const void* qkv, const void* o, const void* dO, const float* lse, void* dqkv, float* dbias_qkv,
                   void* workspace, int B, int N, int H, float scale, void* stream) {
  return passt_attn_bwd_ex(qkv, o, dO, lse, dqkv, dbias_qkv, workspace, B, N, H, scale, 0, stream);
}

int passt_attn_bwd_ex(const void* qkv, const void* o, const void* dO, const float* lse, void* dqkv, float* dbias_qkv,
                      void* workspace, int B, int N, int H, float scale, int flags, void* stream) {
  using namespace pb;
  if (B <= 0 || N <= 0 || H <= 0 || !workspace) return PB_ERR_BAD_ARG;
  if (dbias_qkv != nullptr && H * kBHd != 768) return PB_ERR_BAD_ARG;   // per-CTA bias partials are sized for C = 768
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int C = H * kBHd;
  float* dq_acc = reinterpret_cast<float*>(workspace);
  float* Dsum = dq_acc + size_t(B) * N * C;
  const int Npad = ((N + kTile - 1) / kTile) * kTile;
  // one memset covers dq_acc and the padded D buffer (pad rows of D must be finite: 0 * NaN would poison dS)
  if (!(flags & 1)) {
    PB_CUDA_TRY(cudaMemsetAsync(dq_acc, 0, size_t(B) * N * C * 4 + size_t(B) * H * Npad * 4, st));
    const long long warps = (long long)B * N;
    PB_LAUNCH(attn_dsum_kernel, int((warps * 32 + 255) / 256), 256, 0, st, (const __nv_bfloat16*)o,
              (const __nv_bfloat16*)dO, Dsum, B, N, H, Npad);
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
  p.n_kvt = (N + kTile - 1) / kTile;
  p.total_items = B * H * p.n_kvt;
  p.Npad = p.n_kvt * kTile;
  p.dq_acc = dq_acc;
  p.dbias = dbias_qkv;
  p.timeline = pb::g_attn_bwd_timeline;
  if (g_attn_bwd_variant < 0) {
    const char* e = getenv("PASST_B200_ATTN_BWD");
    g_attn_bwd_variant = (e != nullptr && e[0] == '2') ? 2 : 1;
  }
  const int grid = p.total_items < g_sm_limit ? p.total_items : g_sm_limit;
  if (g_attn_bwd_variant == 2) {
    PB_SET_SMEM_ONCE(AttnBwdSmem::kTotal + kSmemAlignSlack, attn_bwd2_kernel);
    PB_LAUNCH(attn_bwd2_kernel, grid, kBwd2Threads, AttnBwdSmem::kTotal + kSmemAlignSlack, st, tmQKV, tmdO, tmdQKV,
              tmdQacc, p);
  } else {
    PB_SET_SMEM_ONCE(AttnBwdSmem::kTotal + kSmemAlignSlack, attn_bwd_kernel);
    PB_LAUNCH(attn_bwd_kernel, grid, kBwdThreads, AttnBwdSmem::kTotal + kSmemAlignSlack, st, tmQKV, tmdO, tmdQKV,
              tmdQacc, p);
  }
  {
    if (C % 256 != 0) return PB_ERR_BAD_ARG;
    const int rows = B * N;
    const int col_blocks = C / 256;
    int row_blocks = (kNumSMs * 4 + col_blocks - 1) / col_blocks;
    int rows_per_cta = (rows + row_blocks - 1) / row_blocks;
    if (rows_per_cta < 8) rows_per_cta = 8;
    row_blocks = (rows + rows_per_cta - 1) / rows_per_cta;
    PB_LAUNCH(attn_dq_pack_kernel, dim3(col_blocks, row_blocks), 256, 0, st, dq_acc, (__nv_bfloat16*)dqkv, dbias_qkv,
              rows, C, scale, rows_per_cta);
  }
  return 0;
}

}  // extern "C"
