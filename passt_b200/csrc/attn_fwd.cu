// passt_b200 — fused multi-head attention forward on tcgen05 / TMEM (sm_100a), head_dim 64, non-causal.
//
// Replaces Attention.forward's q@k^T * scale -> softmax -> @v (reference models/passt.py:345-358) without ever
// materialising the [B, H, N, N] score tensor.  Input is the packed qkv GEMM output [B, N, 3*H*64] (bf16) exactly
// as nn.Linear(dim, 3*dim) lays it out (:345 reshape(B, N, 3, H, hd)); output is [B, N, H*64] (bf16), i.e. the
// (attn @ v).transpose(1, 2).reshape(B, N, C) layout (:358), plus the per-row log-sum-exp for the backward pass.
//
// Persistent CTAs walk a strided list of (128-query tile, head, clip) items.  Warp roles:
//   warp 0    : TMA producer (Q tiles and K/V tiles through 2-stage rings, running ahead across items)
//   warp 1    : TMEM allocator + tcgen05.mma issuer (descriptors precomputed; an elected lane only issues)
//   warps 2-5 : softmax, one thread per query row (tcgen05.ld 32x32b)
//   S   = Q K^T  : A = Q tile (smem, K-major), B = K tile (smem, K-major)              -> TMEM [0,128)
//   O  += P_j V_j: A = P_j in TENSOR MEMORY (bf16 pairs written by the softmax threads) -> TMEM [192,256),
//                  B = V tile (smem, MN-major view of the [keys, hd] tile); O accumulates in TMEM across key tiles
// Softmax is single-pass with a lagged reference maximum (FA4-style): probabilities of tile j are taken relative to
// the maximum known before the tile; the running sum and the TMEM-resident O are rescaled only when the maximum
// grows by more than 2^8, and a guarded slow path redoes a tile whose scores exceed the reference by more than 2^64
// (exactness is preserved in every case; bf16/fp32 have the exponent range for the lag).
// TMEM columns: S [0,128)  P [128,192)  O [192,256) -> 256 allocated, two CTAs co-reside per SM.
#include "common.cuh"
#include <cstdlib>

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
  float* lse;        // [B, H, Npad] log2-domain LSE of the scaled scores (Npad = 128*ceil(N/128); pad rows = +inf)
  long long* timeline;  // bring-up only: clock64 stamps of CTA 0 (nullptr in production)
};
#define PB_STAMP(role, idx)                                                                     \
  do {                                                                                           \
    if (p.timeline != nullptr && blockIdx.x == 0 && (idx) < 512) p.timeline[(role) * 512 + (idx)] = clock64(); \
  } while (0)

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
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
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
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (p.N + kKvTile - 1) / kKvTile;
  // the last key tile is trimmed to whole 32-column chunks: its S MMA runs with N = cols_last, the softmax touches
  // cols_last / 32 chunks and P V contracts over cols_last keys (N = 474: 96 instead of 128 -> 6 % less exp2 / MMA work)
  const int cols_last = ((p.N - (n_kv - 1) * kKvTile + 31) / 32) * 32;

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
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_gate();
  const uint32_t tmem_S = tmem_base;         // fp32 scores, 128 columns
  const uint32_t tmem_P = tmem_base + 128;   // bf16 probabilities (pairs), 64 columns
  const uint32_t tmem_O = tmem_base + 192;   // fp32 output accumulator, 64 columns
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
          PB_STAMP(2, t);
          uint8_t* sK = sKV + s * (2 * kKvTile * kHd * 2);
          uint8_t* sV = sK + kKvTile * kHd * 2;
          mbar_arrive_expect_tx(&kv_full[s], 2 * kKvTile * kHd * 2);
          tma_load_3d(sK, &tmQKV, &kv_full[s], C + h * kHd, j * kKvTile, b);
          tma_load_3d(sV, &tmQKV, &kv_full[s], 2 * C + h * kHd, j * kKvTile, b);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // S = Q K^T : A, B K-major (smem)
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);    // O = P V   : A TMEM, B (V) MN-major
    // all shared-memory descriptors are computed once, warp-uniformly, outside the critical path; inside the loop an
    // elected lane only issues the tcgen05 instructions
    uint64_t dQ[2][4], dK[2][4], dV[2][8];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const uint32_t aQ = smem_u32(sQ + s * (kQTile * kHd * 2));
      const uint32_t aK = smem_u32(sKV + s * (2 * kKvTile * kHd * 2));
      const uint32_t aV = aK + kKvTile * kHd * 2;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dQ[s][k] = make_smem_desc_sw128(aQ + k * 32, 16, 1024);
        dK[s][k] = make_smem_desc_sw128(aK + k * 32, 16, 1024);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) dV[s][k] = make_smem_desc_sw128(aV + k * 2048, 8192, 1024);
    }
    const uint32_t idesc_s_last = make_idesc_bf16(128, uint32_t(cols_last), 0, 0);
    auto issue_s = [&](uint32_t t, uint32_t qs, bool last_tile) {
      const uint32_t s = t & 1;
      const uint32_t idesc = last_tile ? idesc_s_last : idesc_s;
      mbar_wait(&kv_full[s], (t >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_S, qs ? dQ[1][k] : dQ[0][k], s ? dK[1][k] : dK[0][k], idesc, k > 0 ? 1u : 0u);
        tc_commit(s_full);
      }
      __syncwarp();
    };
    uint32_t t = 0, n = 0;
    for (int it = blockIdx.x; it < p.total_items; it += gridDim.x, ++n) {
      const uint32_t qs = n & 1;
      const bool has_next = (it + int(gridDim.x) < p.total_items);
      mbar_wait(&q_full[qs], (n >> 1) & 1);
      if (n == 0) issue_s(t, qs, n_kv == 1);   // later items: S of their first tile was issued ahead (below)
      for (int j = 0; j < n_kv; ++j, ++t) {
        const uint32_t s = t & 1;
        mbar_wait(p_full, t & 1);   // P_t is in TMEM and S_t has been consumed
        tc_fence_after();
        if (lane == 0) PB_STAMP(1, t * 3 + 0);
        // next scores first (the softmax warps can start on them while PV_t runs)
        if (j + 1 < n_kv) {
          issue_s(t + 1, qs, j + 2 == n_kv);
        } else if (has_next) {
          mbar_wait(&q_full[qs ^ 1], ((n + 1) >> 1) & 1);
          issue_s(t + 1, qs ^ 1, n_kv == 1);
        }
        if (lane == 0) PB_STAMP(1, t * 3 + 1);
        const int n_pv = (j + 1 == n_kv) ? cols_last / 16 : 8;   // 16 keys per MMA
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < n_pv)
              umma_bf16_ts(tmem_O, tmem_P + k * 8, s ? dV[1][k] : dV[0][k], idesc_o, (j > 0 || k > 0) ? 1u : 0u);
          tc_commit(o_full);
          tc_commit(&kv_empty[s]);
        }
        __syncwarp();
        if (lane == 0) PB_STAMP(1, t * 3 + 2);
      }
    }
  } else {
    // ===================== softmax / accumulate warps =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;            // query row inside the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    constexpr float kRescaleTh = 8.0f;     // lagged-max: rescale O / l only when the maximum grew by > 2^8
    constexpr float kGuardTh = 64.0f;      // redo a tile whose scores exceed the reference by more than 2^64
    uint32_t t = 0, n = 0;
    for (int it = blockIdx.x; it < p.total_items; it += gridDim.x, ++n) {
      const int qt = it % p.n_qt, h = (it / p.n_qt) % p.H, b = it / (p.n_qt * p.H);
      const int q0 = qt * kQTile;
      float m_used = 0.f, m_seen = 0.f, l_run = 0.f;     // scaled (log2) units

      for (int j = 0; j < n_kv; ++j, ++t) {
        const int kv_valid = min(kKvTile, p.N - j * kKvTile);
        const bool full_tile = (kv_valid == kKvTile);     // warp-uniform: only the last key tile needs masking
        const int n_chunks = (j + 1 == n_kv) ? cols_last / 32 : 4;
        mbar_wait(s_full, t & 1);
        tc_fence_after();
        if (warp == 2 && lane == 0) PB_STAMP(0, t * 5 + 0);
        if (j == 0) {
          // reference maximum of a new row: the first 32 scores (key 0 is always valid)
          uint32_t v[32];
          tmem_ld_x32(tmem_S + lane_addr, v);
          tmem_ld_wait();
          float mx = __uint_as_float(v[0]);
#pragma unroll
          for (int i = 1; i < 32; ++i)
            if (i < kv_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
          m_used = mx * p.scale_log2;
          m_seen = m_used;
        } else if (__any_sync(0xffffffffu, m_seen - m_used > kRescaleTh)) {
          // rare: bring l and the TMEM accumulator to the new reference (PV_{t-1} must have landed)
          const float alpha = ex2_approx(m_used - m_seen);     // == 1 for rows whose maximum did not move
          mbar_wait(o_full, (t - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_x32(tmem_O + lane_addr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_x32(tmem_O + lane_addr + c * 32, v);
          }
          tmem_st_wait();
          l_run *= alpha;
          m_used = m_seen;
        }
        if (warp == 2 && lane == 0) PB_STAMP(0, t * 5 + 1);
        // single pass: P = exp2(S*c - m_used) -> bf16 pairs -> TMEM; track the tile maximum on the side
        bool redo = false;
        do {
          float rs0 = 0.f, rs1 = 0.f, mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c >= n_chunks) break;            // trimmed last tile (warp-uniform)
            uint32_t v[32];
            tmem_ld_x32(tmem_S + lane_addr + c * 32, v);
            tmem_ld_wait();
            uint32_t pk[16];
            if (full_tile) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
                mx0 = fmaxf(mx0, s0);
                mx1 = fmaxf(mx1, s1);
                const float e0 = ex2_approx(fmaf(s0, p.scale_log2, -m_used));
                const float e1 = ex2_approx(fmaf(s1, p.scale_log2, -m_used));
                rs0 += e0;
                rs1 += e1;
                pk[i >> 1] = pack_bf16(e0, e1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const bool ok0 = c * 32 + i < kv_valid, ok1 = c * 32 + i + 1 < kv_valid;
                const float s0 = ok0 ? __uint_as_float(v[i]) : -INFINITY;
                const float s1 = ok1 ? __uint_as_float(v[i + 1]) : -INFINITY;
                mx0 = fmaxf(mx0, s0);
                mx1 = fmaxf(mx1, s1);
                const float e0 = ex2_approx(fmaf(s0, p.scale_log2, -m_used));   // exp2(-inf) = 0 for masked keys
                const float e1 = ex2_approx(fmaf(s1, p.scale_log2, -m_used));
                rs0 += e0;
                rs1 += e1;
                pk[i >> 1] = pack_bf16(e0, e1);
              }
            }
            tmem_st_x16(tmem_P + lane_addr + c * 16, pk);
          }
          const float tile_max = fmaxf(mx0, mx1) * p.scale_log2;
          redo = false;
          if (__any_sync(0xffffffffu, tile_max - m_used > kGuardTh)) {
            // guarded slow path (scores far above the reference): move the reference and redo this tile; S is intact
            const float m_new = fmaxf(m_used, tile_max);
            const float alpha = ex2_approx(m_used - m_new);
            if (j > 0) {
              mbar_wait(o_full, (t - 1) & 1);
              tc_fence_after();
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld_x32(tmem_O + lane_addr + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                tmem_st_x32(tmem_O + lane_addr + c * 32, v);
              }
              tmem_st_wait();
            }
            l_run *= alpha;
            m_used = m_new;
            m_seen = fmaxf(m_seen, m_new);
            redo = true;
          } else {
            l_run += rs0 + rs1;
            m_seen = fmaxf(m_seen, tile_max);
          }
        } while (redo);
        if (warp == 2 && lane == 0) PB_STAMP(0, t * 5 + 3);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full);
        if (warp == 2 && lane == 0) PB_STAMP(0, t * 5 + 4);
      }
      // ---- item epilogue: O (TMEM) / l -> bf16 -> swizzled staging (this item's Q buffer) -> TMA store
      mbar_wait(o_full, (t - 1) & 1);
      tc_fence_after();
      const float inv_l = 1.0f / l_run;
      // log2-domain LSE, padded rows get +inf (the backward pass turns them into P = 0)
      p.lse[(size_t(b) * p.H + h) * (size_t(p.n_qt) * kQTile) + q0 + r] =
          (q0 + r < p.N) ? (m_used + log2f(l_run)) : INFINITY;
      uint8_t* stage = sQ + (n & 1) * (kQTile * kHd * 2) + q * 4096;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tmem_O + lane_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int ch = c * 4 + g;
          uint4 o;
          o.x = pack_bf16(__uint_as_float(v[g * 8 + 0]) * inv_l, __uint_as_float(v[g * 8 + 1]) * inv_l);
          o.y = pack_bf16(__uint_as_float(v[g * 8 + 2]) * inv_l, __uint_as_float(v[g * 8 + 3]) * inv_l);
          o.z = pack_bf16(__uint_as_float(v[g * 8 + 4]) * inv_l, __uint_as_float(v[g * 8 + 5]) * inv_l);
          o.w = pack_bf16(__uint_as_float(v[g * 8 + 6]) * inv_l, __uint_as_float(v[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(stage + lane * 128 + ((ch ^ (lane & 7)) << 4)) = o;
        }
      }
      // O has been read: the next item's first PV (accumulate = 0) may overwrite it.  That MMA needs this thread's
      // p_full arrival for the next tile, which comes later in program order, so no extra barrier is required.
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

long long* g_attn_timeline = nullptr;

// attn_fwd2.cu: ping-pong variant (one CTA per SM, two query tiles in flight)
int launch_attn_fwd2(const void* qkv, void* out, float* lse, int B, int N, int H, float scale, bool early_s,
                     cudaStream_t st);
// attn_fwd3.cu: ping-pong with two softmax threads per query row (16 softmax warps)
int launch_attn_fwd3(const void* qkv, void* out, float* lse, int B, int N, int H, float scale, cudaStream_t st);
int g_attn_fwd_variant = [] {
  // 2 (default): ping-pong; 4: ping-pong, next Q K^T issued as soon as the scores are in registers; 3: ping-pong, 2 threads
  // per row; 1: 2 CTAs per SM
  const char* e = getenv("PASST_B200_ATTN_FWD");
  return (e != nullptr && e[0] == '1') ? 1 : (e != nullptr && e[0] == '3') ? 3 : (e != nullptr && e[0] == '4') ? 4 : 2;
}();

}  // namespace pb

extern "C" {
// bring-up: device buffer of >= 3*512 int64 receiving clock64 stamps of CTA 0 (NULL disables)
void passt_attn_debug_timeline(void* buf) { pb::g_attn_timeline = reinterpret_cast<long long*>(buf); }

// 2 (default): ping-pong kernel (attn_fwd2.cu); 1: the two-CTAs-per-SM kernel of this file
void passt_attn_fwd_set_variant(int v) { pb::g_attn_fwd_variant = (v == 1 || v == 3 || v == 4) ? v : 2; }


// qkv: bf16 [B, N, 3*H*64]; out: bf16 [B, N, H*64]; lse: fp32 [B, H, Npad], Npad = 128*ceil(N/128), log2 domain
int passt_attn_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, float scale, void* stream) {
  using namespace pb;
  if (B <= 0 || N <= 0 || H <= 0) return PB_ERR_BAD_ARG;
  if ((g_attn_fwd_variant == 2 || g_attn_fwd_variant == 4) && g_attn_timeline == nullptr)
    return launch_attn_fwd2(qkv, out, lse, B, N, H, scale, g_attn_fwd_variant == 4, reinterpret_cast<cudaStream_t>(stream));
  if (g_attn_fwd_variant == 3 && g_attn_timeline == nullptr)
    return launch_attn_fwd3(qkv, out, lse, B, N, H, scale, reinterpret_cast<cudaStream_t>(stream));
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
  p.timeline = pb::g_attn_timeline;
  PB_SET_SMEM_ONCE(AttnFwdSmem::kTotal + kSmemAlignSlack, attn_fwd_kernel);
  const int grid = p.total_items < 2 * g_sm_limit ? p.total_items : 2 * g_sm_limit;
  PB_LAUNCH(attn_fwd_kernel, grid, kAttnThreads, AttnFwdSmem::kTotal + kSmemAlignSlack, reinterpret_cast<cudaStream_t>(stream), tmQKV, tmO, p);
  return 0;
}

}  // extern "C"
