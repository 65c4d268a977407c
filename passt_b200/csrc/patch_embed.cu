// passt_b200 — single-kernel patch embedding: TMA gather of the KEPT 16x16 patches -> bf16 operand tile in shared memory
// -> tcgen05 GEMM against the conv weight -> + (bias, time/freq positional embedding, cls/dist rows) -> fp32 tokens.
//
// Replaces PatchEmbed.proj (Conv2d(1, 768, 16, stride), models/passt.py:315,323), the positional adds (:527-529), the
// Patchout gathers (:535-552: dropped patches are never read) and the cls/dist token assembly (:557-564), and -- when
// the caller folds spectrogram mixup in (ex_audioset.py:173-177) -- the mix of the two source clips.  The patch rows
// never exist in HBM: HBM traffic is the kept patches of the mel (read once, through TMA boxes) and the token tensor.
//
// Persistent CTAs walk 128-token tiles.  Warp roles (480 threads):
//   warp 0     : strip producer.  Kept patches are fetched in STRIPS: consecutive tokens of a tile that lie in the same clip
//                and patch row and whose columns fit a 160-frame window share ONE 3-D box [1 clip, 16 mel bins, 160 frames]
//                (two boxes with mixup) -- up to 15 patches per box at stride 10 -- into a 3-deep staging ring.
//                TMA tile loads need a 16-byte aligned start in the contiguous dimension (measured: a box starting at
//                frame 10 raises an illegal-instruction fault, frames 0/4/8/12 work; tests/probe/tma_probe.cu): a strip
//                starts at its first patch's frame rounded down to a multiple of 4; the converter applies the offsets.
//   warp 10    : weight producer: conv-weight k-blocks [256 out x 64 k] into a 2-deep ring
//   warp 1     : tcgen05.mma issuer (M = 128 tokens, N = 256 channels, K = 256 taps; 3 channel tiles per token tile,
//                accumulators double-buffered in TMEM)
//   warps 2-5  : converter: staged fp32 patches (x lam + partner x (1 - lam)) -> bf16 -> K-major SWIZZLE_128B A tile
//   warps 6-9  : epilogue: tcgen05.ld + TMA-fetched token-table chunk -> in-place sum in shared memory
//   warps 11-14: one helper lane per epilogue warp: token-table chunk requests and TMA stores of the sums (fp32 tokens)
#include "common.cuh"

#include <type_traits>

namespace pb {

constexpr int kPeThreads = 480;
constexpr int kPeDm = 768;
constexpr int kPeMaxTok = 16;                // patches per strip (8 converter threads each)
constexpr int kPeSW = 160;                   // frames per strip box
constexpr int kPeSrc = 16 * kPeSW * 4;       // 10240 B per source clip and strip (16 mel rows x 160 frames, fp32)
constexpr int kPeBuf = 2 * kPeSrc;           // 20480 B per staging slot (two sources)
constexpr int kPeSlots = 3;                  // staging ring depth with mixup (two sources per slot); 6 single-source slots without
constexpr int kPeEpiSlots = 4;               // per epilogue warp: ring of [32 rows x 16 cols] fp32 buffers
constexpr int kPeEpiBuf = 32 * 64;

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct PatchEmbedParams {
  const float* tab;          // [ntok, 768] additive token table (also described by tmTab)
  const int* patch_f;        // [ntok - 2]
  const int* patch_t;
  const int* mix_perm;       // [B] or nullptr
  const float* mix_lam;      // [B]
  int B, ntok, M, m_tiles, fstride, tstride;
};

struct PatchEmbedSmem {
  static constexpr int kA = 0;                              // 4 k-block atoms x [128 rows x 128 B] = 64 KB
  static constexpr int kB = kA + 65536;                     // 2 stages x [256 n x 64 k] bf16 = 64 KB
  static constexpr int kStage = kB + 65536;                 // 3 slots x 2 sources x 10240 B = 60 KB
  static constexpr int kEpi = kStage + kPeSlots * kPeBuf;   // 4 epilogue warps x 4 x 2 KB = 32 KB
  static constexpr int kMeta = kEpi + 4 * kPeEpiSlots * kPeEpiBuf;   // per staging slot: {first row, patches, start frame, last}
  static constexpr int kBars = kMeta + 128;
  static constexpr int kTotal = kBars + 512;
};

__global__ void __launch_bounds__(kPeThreads, 1)
patch_embed_kernel(const __grid_constant__ CUtensorMap tmMel, const __grid_constant__ CUtensorMap tmW,
                   const __grid_constant__ CUtensorMap tmTab, const __grid_constant__ CUtensorMap tmOut,
                   const PatchEmbedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* sA = smem + PatchEmbedSmem::kA;
  uint8_t* sB = smem + PatchEmbedSmem::kB;
  uint8_t* sStage = smem + PatchEmbedSmem::kStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PatchEmbedSmem::kBars);
  uint64_t* st_full = bars;            // [6] staging strip landed (TMA tx)
  uint64_t* st_empty = bars + 6;       // [6] 4 arrivals (converter warps)
  uint64_t* b_full = bars + 12;        // [2]
  uint64_t* b_empty = bars + 14;       // [2] tcgen05.commit
  uint64_t* a_full = bars + 16;        // [1] 4 arrivals: A tile of this token tile is complete
  uint64_t* a_empty = bars + 17;       // [1] commit after the tile's last MMA
  uint64_t* t_full = bars + 18;        // [2] accumulator ready
  uint64_t* t_empty = bars + 20;       // [2] 4 arrivals (epilogue warps)
  uint64_t* e_full = bars + 22;        // [4 warps][4 slots] token-table chunk landed (TMA tx)
  uint64_t* e_ready = bars + 38;       // [4 warps][4 slots] sum written into the slot (1 arrival, epilogue lane 0)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 54);
  int4* s_meta = reinterpret_cast<int4*>(smem + PatchEmbedSmem::kMeta);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool mixing = (p.mix_perm != nullptr);
  // staging ring: 3 slots of two sources with mixup, 6 slots of one source without (strip latency ~2 us: depth matters)
  const uint32_t nslots = mixing ? uint32_t(kPeSlots) : uint32_t(2 * kPeSlots);
  const uint32_t slot_bytes = mixing ? uint32_t(kPeBuf) : uint32_t(kPeSrc);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmMel);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmTab);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < 2 * kPeSlots; ++s) { mbar_init(&st_full[s], 1); mbar_init(&st_empty[s], 4); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1);
      mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4);
    }
    for (int s = 0; s < 16; ++s) { mbar_init(&e_full[s], 1); mbar_init(&e_ready[s], 1); }
    mbar_init(a_full, 4);
    mbar_init(a_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_gate();

  if (warp == 0) {
    // ===================== strip producer =====================
    // The whole warp scans the kept-patch list (lane l looks at the token l places ahead: one round of loads and a ballot
    // per strip -- a single lane chasing patch_f / patch_t entry by entry took ~1 us per strip and starved the converter);
    // lane 0 issues the boxes.
    uint32_t s = 0, ph = 0;        // staging slot and its phase
    // emit one staging strip: cnt consecutive token rows starting at tile row `first` (cnt == 0: terminator)
    auto emit = [&](int first, int cnt, int ws, int f0, int b, int last) {
      if (lane == 0) {
        mbar_wait(&st_empty[s], ph ^ 1);
        s_meta[s] = make_int4(first, cnt, ws, last);
        if (cnt == 0) {
          mbar_arrive(&st_full[s]);
        } else {
          uint8_t* dst = sStage + s * slot_bytes;
          mbar_arrive_expect_tx(&st_full[s], slot_bytes);
          tma_load_3d(dst, &tmMel, &st_full[s], ws, f0, b);
          if (mixing) tma_load_3d(dst + kPeSrc, &tmMel, &st_full[s], ws, f0, __ldg(p.mix_perm + b));
        }
      }
      if (++s == nslots) { s = 0; ph ^= 1; }
    };
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
      const int row0 = mt * 128, row_end = min(p.M, row0 + 128);
      int r = row0;
      int b = r / p.ntok, n = r - b * p.ntok;
      while (r < row_end) {
        if (n < 2) { ++r; ++n; continue; }                 // cls / dist rows carry no patch
        const int nl = n + lane;
        const bool in = lane < kPeMaxTok && r + lane < row_end && nl < p.ntok;
        const int fl = in ? __ldg(p.patch_f + nl - 2) : -1;
        const int tl = in ? __ldg(p.patch_t + nl - 2) * p.tstride : 0;
        const int pf0 = __shfl_sync(0xffffffffu, fl, 0), t0 = __shfl_sync(0xffffffffu, tl, 0);
        const int ws = t0 & ~3;
        const bool ok = in && fl == pf0 && tl >= t0 && tl + 16 <= ws + kPeSW;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        const int cnt = __ffs(~m) - 1;                     // leading run of patches that share the strip (>= 1)
        emit(r - row0, cnt, ws, pf0 * p.fstride, b, 0);
        r += cnt; n += cnt;
        if (n >= p.ntok) { n -= p.ntok; ++b; }
      }
      emit(0, 0, 0, 0, 0, 1);                              // terminator: the converter closes the tile on it
    }
  } else if (warp == 10) {
    // ===================== weight producer: 12 stages [256 out x 64 k] per token tile =====================
    if (lane == 0) {
      uint32_t bb = 0;
      for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
        for (int wq = 0; wq < 12; ++wq, ++bb) {
          const uint32_t s = bb & 1;
          mbar_wait(&b_empty[s], ((bb >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&b_full[s], 256 * 64 * 2);
          tma_load_2d(sB + s * 32768, &tmW, &b_full[s], (wq & 3) * 64, (wq >> 2) * 256);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(128, 256, 0, 0);
    const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
    uint32_t bb = 0, acc = 0, tiles = 0;
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, ++tiles) {
      mbar_wait(a_full, tiles & 1);
      tc_fence_after();
      for (int nt = 0; nt < 3; ++nt, ++acc) {
        const uint32_t as = acc & 1;
        mbar_wait(&t_empty[as], ((acc >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * 256;
        for (int kb = 0; kb < 4; ++kb, ++bb) {
          const uint32_t s = bb & 1;
          mbar_wait(&b_full[s], (bb >> 1) & 1);
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(aA + kb * 16384, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(aB + s * 32768, 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ss(tmem_d, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit(&b_empty[s]);
            if (kb == 3) {
              tc_commit(&t_full[as]);
              if (nt == 2) tc_commit(a_empty);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 6) {
    // ===================== converter: staging (fp32, 64 B per patch row) -> bf16 A tile =====================
    const int ct = threadIdx.x - 64;          // 0..127
    const int pi = ct >> 3;                   // patch inside the strip (0..15)
    const int part = ct & 7;                  // this thread converts patch columns kx = 2*part, 2*part + 1
    uint32_t s = 0, ph = 0, tiles = 0;
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, ++tiles) {
      mbar_wait(a_empty, (tiles & 1) ^ 1);    // the previous tile's MMAs have finished reading A
      {
        // rows without a patch (cls / dist tokens, rows past M) are zero in the operand tile: thread ct owns tile row ct
        const int row = mt * 128 + ct;
        if (row >= p.M || (row % p.ntok) < 2) {
#pragma unroll
          for (int kb = 0; kb < 4; ++kb)
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(sA + kb * 16384 + ct * 128 + (c << 4)) = make_uint4(0, 0, 0, 0);
        }
      }
      for (;;) {
        mbar_wait(&st_full[s], ph);
        const int4 meta = s_meta[s];          // {first tile row, patches, strip start frame, last}
        if (pi < meta.y) {
          const int arow = meta.x + pi;
          const int row = mt * 128 + arow;
          const int b = row / p.ntok, n = row - b * p.ntok;
          const float lam = mixing ? __ldg(p.mix_lam + b) : 1.f;
          const float* src =
              reinterpret_cast<const float*>(sStage + s * slot_bytes) + (__ldg(p.patch_t + n - 2) * p.tstride - meta.z);
          // the 8 threads of a patch split its 16 columns (2 each) and walk the 16 rows: neighbouring lanes read
          // neighbouring frames (bank-conflict free; splitting by rows would put all 8 on one bank, row pitch 160 floats)
          const int kx = 2 * part;
#pragma unroll
          for (int ky = 0; ky < 16; ++ky) {
            float a0 = src[ky * kPeSW + kx], a1 = src[ky * kPeSW + kx + 1];
            if (mixing) {
              a0 = a0 * lam + src[kPeSrc / 4 + ky * kPeSW + kx] * (1.0f - lam);
              a1 = a1 * lam + src[kPeSrc / 4 + ky * kPeSW + kx + 1] * (1.0f - lam);
            }
            // k = ky*16 + kx: k-block atom ky/4, 16-byte chunk (ky%4)*2 + kx/8 inside the 128-byte row (XOR-swizzled),
            // byte (kx%8)*2 inside the chunk
            const int c = (ky & 3) * 2 + (kx >> 3);
            *reinterpret_cast<uint32_t*>(sA + (ky >> 2) * 16384 + arow * 128 + ((c ^ (arow & 7)) << 4) + (kx & 7) * 2) =
                pack_bf16(a0, a1);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&st_empty[s]);    // this warp has read its share of the slot (and its metadata)
        if (++s == nslots) { s = 0; ph ^= 1; }
        if (meta.w) break;
      }
      fence_proxy_async();                           // A tile written by generic stores, read by the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
    }
  } else if (warp < 10) {
    // ===================== epilogue (warps 6-9) =====================
    // Per [32 rows x 16 cols] chunk: the token-table chunk arrives by TMA (SWIZZLE_64B box, requested by this warp's
    // helper up to three chunks ahead through a 4-slot ring), the accumulator chunk by tcgen05.ld; the sum overwrites the
    // slot in place and the helper stores it by TMA.  Neither the table read nor the token store costs LSU line requests
    // (a lane walking its own 3 KB row costs 32 per instruction; a padded transposition with coalesced accesses was
    // slower still), and the TMA bookkeeping is off this warp: one warp per scheduler exposes every instruction's
    // latency (~1000 clk per chunk with the bookkeeping inline).  Measurements: profiles/r2_patch_embed_iterations.txt.
    const int ew = warp - 6;
    const int q = warp & 3;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t swz = uint32_t((lane >> 1) & 3);
    const uint32_t ring = smem_u32(smem + PatchEmbedSmem::kEpi + ew * (kPeEpiSlots * kPeEpiBuf));
    uint64_t* my_full = e_full + ew * kPeEpiSlots;
    uint64_t* my_ready = e_ready + ew * kPeEpiSlots;
    uint32_t off[4];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) off[ch] = uint32_t(lane) * 64 + ((uint32_t(ch) ^ swz) << 4);
    const int tile_step = (gridDim.x * 128) % p.ntok;
    int n0 = (blockIdx.x * 128 + q * 32) % p.ntok;     // token index of this warp's first row
    uint32_t acc = 0, cs = 0;
    // one accumulator tile (16 chunks); kWrap: this warp's rows cross a clip boundary (token index wraps to 0) -- the TMA
    // box zero-fills the rows past ntok and those lanes read their table row with plain loads, one chunk ahead
    auto run_ntile = [&](auto wrap_tag, int nt, uint32_t taddr, const float* wrow, bool wrapped) {
      constexpr bool kWrap = decltype(wrap_tag)::value;
      float4 wnext[4];
      if (kWrap) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          wnext[i] = wrapped ? __ldg(reinterpret_cast<const float4*>(wrow + nt * 256) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll 1
      for (int c = 0; c < 16; ++c, ++cs) {
        uint32_t v[16];
        tmem_ld_x16(taddr + c * 16, v);
        float4 wcur[4];
        if (kWrap) {
#pragma unroll
          for (int i = 0; i < 4; ++i) wcur[i] = wnext[i];
          if (wrapped && c < 15) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              wnext[i] = __ldg(reinterpret_cast<const float4*>(wrow + nt * 256 + (c + 1) * 16) + i);
          }
        }
        const uint32_t s0 = cs & 3;
        const uint32_t b0 = ring + s0 * kPeEpiBuf;
        mbar_wait(&my_full[s0], (cs >> 2) & 1);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float4 t4 = lds128(b0 + off[ch]);
          if (kWrap) { if (wrapped) t4 = wcur[ch]; }
          float4 o;
          o.x = __uint_as_float(v[4 * ch]) + t4.x; o.y = __uint_as_float(v[4 * ch + 1]) + t4.y;
          o.z = __uint_as_float(v[4 * ch + 2]) + t4.z; o.w = __uint_as_float(v[4 * ch + 3]) + t4.w;
          sts128(b0 + off[ch], o);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&my_ready[s0]);
      }
    };
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
      const bool wrap = n0 + 32 > p.ntok;                 // warp-uniform
      const bool wrapped = n0 + lane >= p.ntok;
      const float* wrow = p.tab + size_t(wrapped ? n0 + lane - p.ntok : 0) * kPeDm;
#pragma unroll 1
      for (int nt = 0; nt < 3; ++nt, ++acc) {
        const uint32_t as = acc & 1;
        mbar_wait(&t_full[as], (acc >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_addr + as * 256;
        if (wrap) run_ntile(std::true_type{}, nt, taddr, wrow, wrapped);
        else run_ntile(std::false_type{}, nt, taddr, wrow, false);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[as]);
      }
      n0 += tile_step;
      if (n0 >= p.ntok) n0 -= p.ntok;
    }
  } else {
    // ===================== epilogue helpers (warps 11-14, one lane each): table-chunk requests and token stores =========
    if (lane == 0) {
      const int ew = warp - 11;
      const int q = (ew + 6) & 3;                          // TMEM quadrant of epilogue warp 6 + ew
      uint8_t* ring = smem + PatchEmbedSmem::kEpi + ew * (kPeEpiSlots * kPeEpiBuf);
      uint64_t* my_full = e_full + ew * kPeEpiSlots;
      uint64_t* my_ready = e_ready + ew * kPeEpiSlots;
      const int tile_step = (gridDim.x * 128) % p.ntok;
      // request iterator: next chunk whose table box has not been requested
      int i_mt = blockIdx.x, i_cc = 0, i_n0 = (blockIdx.x * 128 + q * 32) % p.ntok;
      uint32_t is = 0, cs = 0;
      auto request = [&]() {
        const uint32_t s = is & 3;
        mbar_arrive_expect_tx(&my_full[s], kPeEpiBuf);
        tma_load_2d(ring + s * kPeEpiBuf, &tmTab, &my_full[s], i_cc * 16, i_n0);   // rows past ntok: zero-filled
        ++is;
        if (++i_cc == 48) {
          i_cc = 0; i_mt += gridDim.x;
          i_n0 += tile_step;
          if (i_n0 >= p.ntok) i_n0 -= p.ntok;
        }
      };
      for (int k = 0; k < kPeEpiSlots && i_mt < p.m_tiles; ++k) request();
      for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x) {
        const int wrow0 = mt * 128 + q * 32;
#pragma unroll 1
        for (int cc = 0; cc < 48; ++cc, ++cs) {
          const uint32_t s0 = cs & 3;
          mbar_wait(&my_ready[s0], (cs >> 2) & 1);
          tma_store_2d(&tmOut, ring + s0 * kPeEpiBuf, cc * 16, wrow0);      // rows past M are clipped by the tensor map
          tma_store_commit();
          // every store but this one has read its slot: chunk cs - 1's slot takes the request for chunk cs + 3
          tma_store_wait_read<1>();
          if (cs >= 1 && i_mt < p.m_tiles) request();
        }
      }
      tma_store_wait<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace pb

extern "C" {

// mel f32 [B, Fm, Tm]; w_bf16 [768, 256] (conv weight, [out, ky*16+kx]); tab f32 [ntok, 768] (passt_token_table);
// out f32 [B*ntok, 768].  Returns PB_ERR_BAD_ARG when the mel rows are not 16-byte aligned (Tm % 4 != 0): TMA cannot
// describe such a tensor and the caller uses passt_im2col + passt_gemm_bf16 instead.
int passt_patch_embed(const float* mel, const void* w_bf16, const float* tab, float* out, const int* patch_f,
                      const int* patch_t, int B, int ntok, int Fm, int Tm, int fstride, int tstride,
                      const int* mix_perm, const float* mix_lam, void* stream) {
  using namespace pb;
  if (!mel || !w_bf16 || !tab || !out || !patch_f || !patch_t || B <= 0 || ntok < 2) return PB_ERR_BAD_ARG;
  if ((Tm % 4) != 0 || Fm < 16 || Tm < kPeSW || (mix_perm == nullptr) != (mix_lam == nullptr)) return PB_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(mel) & 15) != 0) return PB_ERR_BAD_ARG;
  if (ntok < 32) return PB_ERR_BAD_ARG;          // the epilogue handles at most one clip boundary per 32 rows
  CUtensorMap tmMel, tmW, tmTab, tmOut;
  int rc;
  if ((rc = make_tmap_3d(&tmMel, mel, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Tm, Fm, B, uint64_t(Tm) * 4,
                         uint64_t(Fm) * Tm * 4, kPeSW, 16, 1, CU_TENSOR_MAP_SWIZZLE_NONE)))
    return rc;
  if ((rc = make_tmap_2d(&tmW, w_bf16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kPeDm, 256, 256 * 2, 256, 64,
                         CU_TENSOR_MAP_SWIZZLE_128B)))
    return rc;
  const int Mrows = B * ntok;
  if ((rc = make_tmap_2d(&tmTab, tab, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ntok, kPeDm, uint64_t(kPeDm) * 4, 32, 16,
                         CU_TENSOR_MAP_SWIZZLE_64B)))
    return rc;
  if ((rc = make_tmap_2d(&tmOut, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Mrows, kPeDm, uint64_t(kPeDm) * 4, 32, 16,
                         CU_TENSOR_MAP_SWIZZLE_64B)))
    return rc;
  PatchEmbedParams p;
  p.tab = tab; p.patch_f = patch_f; p.patch_t = patch_t; p.mix_perm = mix_perm; p.mix_lam = mix_lam;
  p.B = B; p.ntok = ntok; p.M = B * ntok; p.m_tiles = (p.M + 127) / 128; p.fstride = fstride; p.tstride = tstride;
  const size_t smem_bytes = size_t(PatchEmbedSmem::kTotal) + kSmemAlignSlack;
  static_assert(PatchEmbedSmem::kTotal + kSmemAlignSlack <= 227 * 1024, "patch embed shared memory");
  PB_SET_SMEM_ONCE(227 * 1024, patch_embed_kernel);
  const int grid = p.m_tiles < g_sm_limit ? p.m_tiles : g_sm_limit;
  PB_LAUNCH(patch_embed_kernel, grid, kPeThreads, smem_bytes, reinterpret_cast<cudaStream_t>(stream), tmMel, tmW, tmTab,
            tmOut, p);
  return 0;
}

}  // extern "C"
