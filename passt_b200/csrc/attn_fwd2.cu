// passt_b200 — fused multi-head attention forward, ping-pong variant (sm_100a, head_dim 64, non-causal).
//
// Same math, inputs and outputs as attn_fwd.cu (reference models/passt.py:345-358), different schedule: ONE CTA per SM
// works on TWO 128-query tiles (A, B) of the same (clip, head) at a time, FlashAttention-4 style.  While the 128
// softmax threads of tile A run their exp2 pass, the tensor core executes P_B V and the next Q_B K^T, and vice versa,
// so the MUFU pipe (the hd = 64 bottleneck: one exp2 per score) never waits for an MMA round trip and every K / V
// tile brought in by TMA feeds two query tiles.
//
// Warp roles (320 threads): warp 0 TMA producer | warp 1 tcgen05.mma issuer | warps 2-5 softmax A | warps 6-9 softmax B
// TMEM (512 columns): S_A [0,128)  S_B [128,256)  O_A [256,320)  O_B [320,384)  P_A [384,448)  P_B [448,512)
// MMA issue order per key tile j:   [P_A(j) ready] S_A(j+1), PV_A(j)   [P_B(j) ready] S_B(j+1), PV_B(j)
// Barriers per side X: s_full[X] (S_X landed), p_full[X] (128 arrivals: P_X written, S_X consumed), pv_done[X]
// (P_X V accumulated: P_X may be overwritten, O_X may be rescaled / read).
#include "common.cuh"
#include <cstdlib>

namespace pb {

constexpr int k2Hd = 64;
constexpr int k2Threads = 320;
constexpr int k2Tile = 128;
constexpr int k2KvStages = 3;

struct AttnFwd2Params {
  int N, H, B;
  int n_qt;          // 128-query tiles per (clip, head)
  int n_pairs;       // ceil(n_qt / 2)
  int total_items;   // B * H * n_pairs
  float scale_log2;
  float* lse;        // [B, H, Npad] log2-domain LSE (Npad = 128 * n_qt; pad rows = +inf)
};

struct AttnFwd2Smem {
  static constexpr int kQ = 0;                                        // [2 item slots][2 sides] x 16 KB (also output staging)
  static constexpr int kKV = kQ + 4 * k2Tile * k2Hd * 2;              // k2KvStages x (K 16 KB + V 16 KB)
  static constexpr int kBars = kKV + k2KvStages * 2 * k2Tile * k2Hd * 2;
  static constexpr int kTotal = kBars + 256;
};

// kPrefetch: the TMEM read of score chunk c+1 is issued before chunk c is processed (two register buffers), so the
// tcgen05.ld round trip is paid once per tile instead of once per 32-column chunk.
// exp2 on the FMA / ALU pipes (Cody-Waite split + degree-4 polynomial, relative error ~2e-6 -- far below the bf16
// rounding of P): used for every 4th score when kPolyExp is set, to take a quarter of the exponentials off the
// MUFU (XU) pipe, which also serves the bf16 packs.  x <= ~+8 here; very negative x flushes to 0 like ex2.approx.ftz.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;                 // 1.5 * 2^23: round-to-nearest integer lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);           // f in [-0.5, 0.5]
  float p = fmaf(f, 9.6181291e-3f, 5.5504109e-2f); // minimax-ish Taylor coefficients of 2^f
  p = fmaf(p, f, 2.4022651e-1f);
  p = fmaf(p, f, 6.9314718e-1f);
  p = fmaf(p, f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));   // * 2^round(x)
}

template <bool kPrefetch, bool kPolyExp, bool kEarlyS>
__global__ void __launch_bounds__(k2Threads, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                 const AttnFwd2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sQ = smem + AttnFwd2Smem::kQ;
  uint8_t* sKV = smem + AttnFwd2Smem::kKV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnFwd2Smem::kBars);
  uint64_t* q_full = bars;             // [2 slots][2 sides]  index slot*2+side
  uint64_t* q_empty = bars + 4;        // [2][2], 4 arrivals each (one per softmax warp of the side)
  uint64_t* kv_full = bars + 8;        // [k2KvStages]
  uint64_t* kv_empty = bars + 12;      // [k2KvStages]
  uint64_t* s_full = bars + 16;        // [2 sides]
  uint64_t* p_full = bars + 18;        // [2 sides] 128 arrivals
  uint64_t* pv_done = bars + 20;       // [2 sides]
  uint64_t* s_used = bars + 22;        // [2 sides] 128 arrivals (kEarlyS): every score of S_X(j) is in registers
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (p.N + k2Tile - 1) / k2Tile;
  const int cols_last = ((p.N - (n_kv - 1) * k2Tile + 31) / 32) * 32;   // last key tile trimmed to whole 32-key chunks
  const int C = p.H * k2Hd;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 4; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 4); }
    for (int s = 0; s < k2KvStages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&s_full[x], 1); mbar_init(&p_full[x], 128); mbar_init(&pv_done[x], 1); mbar_init(&s_used[x], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_gate();

  // item -> (pair, head, clip); side X of pair pr is query tile 2*pr + X (side B may not exist for odd n_qt)
  auto item_coords = [&](int it, int& pr, int& h, int& b) {
    pr = it % p.n_pairs;
    h = (it / p.n_pairs) % p.H;
    b = it / (p.n_pairs * p.H);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t t = 0, n = 0;        // t: running key-tile counter (kv ring), n: running item counter (Q slots)
      uint32_t qpar = 0;            // bit qi: parity of the NEXT fill of Q buffer qi (a side may sit out items: odd n_qt)
      for (int it = blockIdx.x; it < p.total_items; it += gridDim.x, ++n) {
        int pr, h, b;
        item_coords(it, pr, h, b);
        const uint32_t slot = n & 1;
        const bool has_b = (2 * pr + 1) < p.n_qt;
        for (int x = 0; x < (has_b ? 2 : 1); ++x) {
          const uint32_t qi = slot * 2 + x;
          mbar_wait(&q_empty[qi], ((qpar >> qi) & 1u) ^ 1u);
          qpar ^= (1u << qi);
          mbar_arrive_expect_tx(&q_full[qi], k2Tile * k2Hd * 2);
          tma_load_3d(sQ + qi * (k2Tile * k2Hd * 2), &tmQKV, &q_full[qi], h * k2Hd, (2 * pr + x) * k2Tile, b);
        }
        for (int j = 0; j < n_kv; ++j, ++t) {
          const uint32_t s = t % k2KvStages;
          mbar_wait(&kv_empty[s], ((t / k2KvStages) & 1) ^ 1);
          uint8_t* sK = sKV + s * (2 * k2Tile * k2Hd * 2);
          uint8_t* sV = sK + k2Tile * k2Hd * 2;
          mbar_arrive_expect_tx(&kv_full[s], 2 * k2Tile * k2Hd * 2);
          tma_load_3d(sK, &tmQKV, &kv_full[s], C + h * k2Hd, j * k2Tile, b);
          tma_load_3d(sV, &tmQKV, &kv_full[s], 2 * C + h * k2Hd, j * k2Tile, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // S = Q K^T : A, B K-major (smem)
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);    // O += P V  : A in TMEM, B (V) MN-major
    const uint32_t idesc_s_last = make_idesc_bf16(128, uint32_t(cols_last), 0, 0);
    const uint32_t aQ0 = smem_u32(sQ), aKV0 = smem_u32(sKV);
    // smem descriptors are rebuilt from warp-uniform bases each time (cheap: one add per field); an elected lane issues
    auto issue_s = [&](uint32_t x, uint32_t slot, uint32_t t, bool last_tile) {
      const uint32_t s = t % k2KvStages;
      mbar_wait(&kv_full[s], (t / k2KvStages) & 1);      // no-op after the first wait on this (stage, phase)
      tc_fence_after();
      const uint32_t aQ = aQ0 + (slot * 2 + x) * (k2Tile * k2Hd * 2);
      const uint32_t aK = aKV0 + s * (2 * k2Tile * k2Hd * 2);
      const uint64_t dq = make_smem_desc_sw128(aQ, 16, 1024), dk = make_smem_desc_sw128(aK, 16, 1024);
      const uint32_t idesc = last_tile ? idesc_s_last : idesc_s;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + x * 128, dq + uint64_t(k * 2), dk + uint64_t(k * 2), idesc, k > 0 ? 1u : 0u);
        tc_commit(&s_full[x]);
      }
      __syncwarp();
    };
    uint32_t t = 0, n = 0;
    uint32_t qpar = 0;               // bit qi: parity of the next q_full wait on Q buffer qi
    auto wait_q = [&](uint32_t qi) {
      mbar_wait(&q_full[qi], (qpar >> qi) & 1u);
      qpar ^= (1u << qi);
    };
    uint32_t tcount[2] = {0, 0};     // per side: running count of key tiles processed (parity of p_full / pv_done)
    for (int it = blockIdx.x; it < p.total_items; it += gridDim.x, ++n) {
      int pr, h, b;
      item_coords(it, pr, h, b);
      const uint32_t slot = n & 1;
      const bool has_b = (2 * pr + 1) < p.n_qt;
      const int nx = has_b ? 2 : 1;
      const bool has_next = (it + int(gridDim.x) < p.total_items);
      bool next_has_b = false;
      if (has_next) {
        int pr2, h2, b2;
        item_coords(it + int(gridDim.x), pr2, h2, b2);
        next_has_b = (2 * pr2 + 1) < p.n_qt;
      }
      if (n == 0) {   // first item of this CTA: nothing was issued ahead
        for (int x = 0; x < nx; ++x) {
          wait_q(slot * 2 + x);
          issue_s(x, slot, t, n_kv == 1);
        }
      }
      for (int j = 0; j < n_kv; ++j, ++t) {
        const uint32_t s = t % k2KvStages;
        const bool last_j = (j + 1 == n_kv);
        for (int x = 0; x < nx; ++x) {
          // next scores of this side first: its softmax threads can start on them while P_X V runs
          auto issue_next_s = [&]() {
            if (!last_j) {
              issue_s(x, slot, t + 1, j + 2 == n_kv);
            } else if (has_next && (x == 0 || next_has_b)) {
              wait_q((slot ^ 1) * 2 + x);
              issue_s(x, slot ^ 1, t + 1, n_kv == 1);
            }
          };
          if (kEarlyS) {
            // S_X(j) has been read into registers (one 32-column chunk of exp2 work is still ahead of the softmax
            // threads): the Q K^T round trip of the next tile runs under that chunk instead of after it
            mbar_wait(&s_used[x], tcount[x] & 1);
            tc_fence_after();
            issue_next_s();
          }
          mbar_wait(&p_full[x], tcount[x] & 1);     // P_X(j) is in TMEM, S_X(j) consumed
          tc_fence_after();
          if (!kEarlyS) issue_next_s();
          const uint32_t aV = aKV0 + s * (2 * k2Tile * k2Hd * 2) + k2Tile * k2Hd * 2;
          const uint64_t dv = make_smem_desc_sw128(aV, 8192, 1024);
          const int n_pv = last_j ? cols_last / 16 : 8;   // 16 keys per MMA
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < n_pv)
                umma_bf16_ts(tmem_base + 256 + x * 64, tmem_base + 384 + x * 64 + k * 8, dv + uint64_t(k * 128), idesc_o,
                             (j > 0 || k > 0) ? 1u : 0u);
            tc_commit(&pv_done[x]);
            if (x == nx - 1) tc_commit(&kv_empty[s]);    // last reader of this K / V stage
          }
          __syncwarp();
          ++tcount[x];
        }
      }
      // a side that exists in the next item but not in this one (odd n_qt) has had no S issued ahead: do it now
      if (has_next && next_has_b && !has_b) {
        wait_q((slot ^ 1) * 2 + 1);
        issue_s(1, slot ^ 1, t, n_kv == 1);
      }
    }
  } else {
    // ===================== softmax warps: side X = (warp - 2) / 4 =====================
    const int x = (warp - 2) >> 2;
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;            // query row inside the tile == TMEM lane
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t tS = tmem_base + x * 128, tO = tmem_base + 256 + x * 64, tP = tmem_base + 384 + x * 64;
    constexpr float kRescaleTh = 8.0f;     // lagged max: rescale O / l only when the maximum grew by > 2^8
    constexpr float kGuardTh = 64.0f;      // redo a tile whose scores exceed the reference by more than 2^64
    uint32_t tc = 0, n = 0;                // tc: key tiles this side has processed (barrier parities)
    for (int it = blockIdx.x; it < p.total_items; it += gridDim.x, ++n) {
      int pr, h, b;
      item_coords(it, pr, h, b);
      const int qt = 2 * pr + x;
      if (qt >= p.n_qt) continue;           // side B of an odd last pair: nothing to do (warp-uniform)
      const int q0 = qt * k2Tile;
      const uint32_t qi = (n & 1) * 2 + x;
      float m_used = 0.f, m_seen = 0.f, l_run = 0.f;     // scaled (log2) units

      auto rescale_o = [&](float alpha) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_x32(tO + lane_addr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_x32(tO + lane_addr + c * 32, v);
        }
        tmem_st_wait();
      };

      for (int j = 0; j < n_kv; ++j, ++tc) {
        const int kv_valid = min(k2Tile, p.N - j * k2Tile);
        const bool full_tile = (kv_valid == k2Tile);
        const int n_chunks = (j + 1 == n_kv) ? cols_last / 32 : 4;
        mbar_wait(&s_full[x], tc & 1);
        tc_fence_after();
        bool pv_waited = false;        // P_X(j-1) V must have completed before P_X is overwritten / O_X touched
        auto wait_prev_pv = [&]() {
          if (!pv_waited && tc > 0) {
            mbar_wait(&pv_done[x], (tc - 1) & 1);
            tc_fence_after();
          }
          pv_waited = true;
        };
        if (j == 0) {
          uint32_t v[32];
          tmem_ld_x32(tS + lane_addr, v);
          tmem_ld_wait();
          float mx = __uint_as_float(v[0]);
#pragma unroll
          for (int i = 1; i < 32; ++i)
            if (i < kv_valid) mx = fmaxf(mx, __uint_as_float(v[i]));
          m_used = mx * p.scale_log2;
          m_seen = m_used;
        } else if (__any_sync(0xffffffffu, m_seen - m_used > kRescaleTh)) {
          const float alpha = ex2_approx(m_used - m_seen);
          wait_prev_pv();
          rescale_o(alpha);
          l_run *= alpha;
          m_used = m_seen;
        }
        bool redo = false;
        do {
          float rs0 = 0.f, rs1 = 0.f, mx0 = -INFINITY, mx1 = -INFINITY;
          uint32_t vv[kPrefetch ? 2 : 1][32];
          if (kPrefetch) tmem_ld_x32(tS + lane_addr, vv[0]);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c >= n_chunks) break;
            if (kPrefetch) {
              tmem_ld_wait();
              if (c + 1 < n_chunks) tmem_ld_x32(tS + lane_addr + (c + 1) * 32, vv[(c + 1) & 1]);
            } else {
              tmem_ld_x32(tS + lane_addr + c * 32, vv[0]);
              tmem_ld_wait();
            }
            uint32_t (&v)[32] = vv[kPrefetch ? (c & 1) : 0];
            if (kEarlyS && c == n_chunks - 1) {
              // every score of this tile is now in registers.  Unless the guard below is going to ask for a second pass
              // over S (same maximum, same test), hand the S columns back to the MMA warp before the last chunk's exp2s
              float cm = -INFINITY;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (full_tile || c * 32 + i < kv_valid) cm = fmaxf(cm, __uint_as_float(v[i]));
              const float tile_max_pre = fmaxf(fmaxf(mx0, mx1), cm) * p.scale_log2;
              if (!__any_sync(0xffffffffu, tile_max_pre - m_used > kGuardTh)) {
                tc_fence_before();
                mbar_arrive(&s_used[x]);
              }
            }
            uint32_t pk[16];
            if (full_tile) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
                mx0 = fmaxf(mx0, s0);
                mx1 = fmaxf(mx1, s1);
                const float e0 = ex2_approx(fmaf(s0, p.scale_log2, -m_used));
                const float e1 = (kPolyExp && (i & 2)) ? exp2_poly(fmaf(s1, p.scale_log2, -m_used))
                                                       : ex2_approx(fmaf(s1, p.scale_log2, -m_used));
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
                const float e0 = ex2_approx(fmaf(s0, p.scale_log2, -m_used));
                const float e1 = ex2_approx(fmaf(s1, p.scale_log2, -m_used));
                rs0 += e0;
                rs1 += e1;
                pk[i >> 1] = pack_bf16(e0, e1);
              }
            }
            if (c == 0) wait_prev_pv();            // the first chunk's math has hidden the wait
            tmem_st_x16(tP + lane_addr + c * 16, pk);
          }
          const float tile_max = fmaxf(mx0, mx1) * p.scale_log2;
          redo = false;
          if (__any_sync(0xffffffffu, tile_max - m_used > kGuardTh)) {
            // guarded slow path (scores far above the reference): move the reference and redo this tile; S is intact
            const float m_new = fmaxf(m_used, tile_max);
            const float alpha = ex2_approx(m_used - m_new);
            if (j > 0) rescale_o(alpha);           // pv_done was waited above
            l_run *= alpha;
            m_used = m_new;
            m_seen = fmaxf(m_seen, m_new);
            redo = true;
          } else {
            l_run += rs0 + rs1;
            m_seen = fmaxf(m_seen, tile_max);
          }
        } while (redo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[x]);
      }
      // ---- item epilogue: O (TMEM) / l -> bf16 -> swizzled staging (this side's Q buffer) -> TMA store
      mbar_wait(&pv_done[x], (tc - 1) & 1);
      tc_fence_after();
      const float inv_l = 1.0f / l_run;
      p.lse[(size_t(b) * p.H + h) * (size_t(p.n_qt) * k2Tile) + q0 + r] =
          (q0 + r < p.N) ? (m_used + log2f(l_run)) : INFINITY;
      uint8_t* stage = sQ + qi * (k2Tile * k2Hd * 2) + q * 4096;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_x32(tO + lane_addr + c * 32, v);
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
      // O_X has been read: the next item's first P_X V (accumulate = 0) needs this thread's next p_full arrival, which
      // comes later in program order, so no extra barrier is required.
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (q0 + q * 32 < p.N) {
          tma_store_3d(&tmO, stage, h * k2Hd, q0 + q * 32, b);
          tma_store_commit();
          tma_store_wait_read<0>();     // the Q buffer may be refilled once the store has read it
        }
        mbar_arrive(&q_empty[qi]);
      }
      __syncwarp();
    }
    if (lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

int launch_attn_fwd2(const void* qkv, void* out, float* lse, int B, int N, int H, float scale, bool early,
                     cudaStream_t st) {
  const int C = H * k2Hd;
  CUtensorMap tmQKV, tmO;
  int rc;
  if ((rc = make_tmap_3d(&tmQKV, qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, 3 * C, N, B, uint64_t(3 * C) * 2,
                         uint64_t(N) * 3 * C * 2, k2Hd, k2Tile, 1, CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  if ((rc = make_tmap_3d(&tmO, out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, C, N, B, uint64_t(C) * 2,
                         uint64_t(N) * C * 2, k2Hd, 32, 1, CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  AttnFwd2Params p;
  p.N = N; p.H = H; p.B = B; p.scale_log2 = scale * 1.4426950408889634f; p.lse = lse;
  p.n_qt = (N + k2Tile - 1) / k2Tile;
  p.n_pairs = (p.n_qt + 1) / 2;
  p.total_items = B * H * p.n_pairs;
  const int grid = p.total_items < g_sm_limit ? p.total_items : g_sm_limit;
  static const bool prefetch = [] {
    const char* e = getenv("PASST_B200_ATTN_PREFETCH");     // default on; 0: one tcgen05.ld round trip per chunk
    return !(e != nullptr && e[0] == '0');
  }();
  static const bool poly = [] {
    const char* e = getenv("PASST_B200_ATTN_POLYEXP");      // default off; 1: every 4th exp2 on the FMA pipe
    return e != nullptr && e[0] == '1';
  }();
#define PB_FWD2(PF, PE, ES)                                                                                   \
  do {                                                                                                        \
    PB_SET_SMEM_ONCE(AttnFwd2Smem::kTotal + kSmemAlignSlack, attn_fwd2_kernel<PF, PE, ES>);                   \
    PB_LAUNCH((attn_fwd2_kernel<PF, PE, ES>), grid, k2Threads, AttnFwd2Smem::kTotal + kSmemAlignSlack, st, tmQKV, tmO, \
              p);                                                                                             \
  } while (0)
  if (early && prefetch) PB_FWD2(true, false, true);
  else if (early) PB_FWD2(false, false, true);
  else if (prefetch && poly) PB_FWD2(true, true, false);
  else if (prefetch) PB_FWD2(true, false, false);
  else if (poly) PB_FWD2(false, true, false);
  else PB_FWD2(false, false, false);
#undef PB_FWD2
  return 0;
}

}  // namespace pb
