// passt_b200 — HBM-bound row kernels around the tensor-core GEMMs (sm_100a).
//
//   ln_fwd          : x_out = x_in (+ delta_bf16); h = LayerNorm(x_out)*gamma+beta  -> bf16, saves mean/rstd
//                     (reference Block.forward residual adds + norm1/norm2, models/passt.py:377-380, :369,:373)
//   ln_bwd          : g_out = g_in + dLN(dh); emits bf16 copy, d_gamma/d_beta and colsum(g_out) (= bias grad of the
//                     linear that produced the residual delta)
//   colsum          : bias gradients of qkv / fc1
//   im2col_patches  : kept 16x16 patches of the mel -> bf16 rows (PatchEmbed.proj as GEMM, passt.py:315,323 +
//                     patchout gathers :535-552: dropped patches are never read)
//   token_table     : per-token additive table = conv bias + time/freq pos-embed (passt.py:527-529) and the
//                     cls/dist rows (passt.py:557-564)
//   token_table_bwd : gradients of the table's sources from the block-0 input gradient
//   cast_transpose  : fp32 master weights -> bf16 [N,K] (fwd / wgrad) and bf16 [K,N] (dgrad)
//   head_fwd/bwd    : final norm on cls/dist rows, average, head LayerNorm + Linear (passt.py:570-588, :463-464)
#include "common.cuh"

namespace pb {

constexpr int D = 768;          // embed_dim of every PaSST arch (models/passt.py:745-912)
constexpr int kVec = D / 128;   // float4 vectors per lane (6)

// ------------------------------------------------------------------------------------------------
// LayerNorm forward (+ fused residual add)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const float* __restrict__ x_in, const __nv_bfloat16* __restrict__ delta, float* __restrict__ x_out,
              __nv_bfloat16* __restrict__ h, float* __restrict__ mean_out, float* __restrict__ rstd_out,
              const float* __restrict__ gamma, const float* __restrict__ beta, int M, float eps) {
  pdl_gate();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const size_t off = size_t(warp) * D;
  float v[kVec * 4];
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 a = *reinterpret_cast<const float4*>(x_in + off + c);
    v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
    if (delta) {
      const uint2 d = *reinterpret_cast<const uint2*>(delta + off + c);
      const __nv_bfloat162 d0 = *reinterpret_cast<const __nv_bfloat162*>(&d.x);
      const __nv_bfloat162 d1 = *reinterpret_cast<const __nv_bfloat162*>(&d.y);
      v[4 * i] += __low2float(d0); v[4 * i + 1] += __high2float(d0);
      v[4 * i + 2] += __low2float(d1); v[4 * i + 3] += __high2float(d1);
    }
  }
  if (x_out) {
#pragma unroll
    for (int i = 0; i < kVec; ++i)
      *reinterpret_cast<float4*>(x_out + off + i * 128 + lane * 4) =
          make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kVec * 4; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kVec * 4; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[warp] = mean;
    if (rstd_out) rstd_out[warp] = rstd;
  }
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
    uint2 o;
    o.x = pack_bf16((v[4 * i] - mean) * rstd * g.x + b.x, (v[4 * i + 1] - mean) * rstd * g.y + b.y);
    o.y = pack_bf16((v[4 * i + 2] - mean) * rstd * g.z + b.z, (v[4 * i + 3] - mean) * rstd * g.w + b.w);
    *reinterpret_cast<uint2*>(h + off + c) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward (+ fused residual-gradient add, bf16 copy, d_gamma/d_beta/colsum partials)
// ------------------------------------------------------------------------------------------------
constexpr int kLnBwdWarps = 8;
constexpr int kLnBwdSmem = kLnBwdWarps * 3 * D * 4;   // per-warp private column accumulators (d_gamma, d_beta, colsum)
// Column accumulators live in shared memory (private per warp, float4 read-modify-write, no atomics) instead of 72
// registers per thread: the kernel is HBM-bound and needs the occupancy more than it needs the registers.
__global__ void __launch_bounds__(kLnBwdWarps * 32, 2)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dh, const float* __restrict__ x, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ gamma, const float* g_in,
              float* g_out, __nv_bfloat16* __restrict__ g_out_bf16, float* __restrict__ dgamma,
              float* __restrict__ dbeta, float* __restrict__ colsum, int M, int rows_per_cta) {
  pdl_gate();
  extern __shared__ float4 ln_acc4[];
  float* acc_all = reinterpret_cast<float*>(ln_acc4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* acc = acc_all + warp * 3 * D;      // [3][D]
  for (int i = lane; i < 3 * D / 4; i += 32) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  for (int r = r0 + warp; r < r1; r += kLnBwdWarps) {
    const size_t off = size_t(r) * D;
    const float mu = mean[r], rs = rstd[r];
    float xh[kVec * 4], dy[kVec * 4];
    float sa = 0.f, sb = 0.f;
    // issue the residual-gradient loads up front so they are in flight during the two row reductions
    float4 gi[kVec];
    if (g_in) {
#pragma unroll
      for (int i = 0; i < kVec; ++i) gi[i] = *reinterpret_cast<const float4*>(g_in + off + i * 128 + lane * 4);
    }
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int c = i * 128 + lane * 4;
      const float4 xv = *reinterpret_cast<const float4*>(x + off + c);
      const uint2 d = *reinterpret_cast<const uint2*>(dh + off + c);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const __nv_bfloat162 d0 = *reinterpret_cast<const __nv_bfloat162*>(&d.x);
      const __nv_bfloat162 d1 = *reinterpret_cast<const __nv_bfloat162*>(&d.y);
      xh[4 * i] = (xv.x - mu) * rs; xh[4 * i + 1] = (xv.y - mu) * rs;
      xh[4 * i + 2] = (xv.z - mu) * rs; xh[4 * i + 3] = (xv.w - mu) * rs;
      dy[4 * i] = __low2float(d0); dy[4 * i + 1] = __high2float(d0);
      dy[4 * i + 2] = __low2float(d1); dy[4 * i + 3] = __high2float(d1);
      float4* a_g = reinterpret_cast<float4*>(acc + c);
      float4* a_b = reinterpret_cast<float4*>(acc + D + c);
      float4 ag = *a_g, ab = *a_b;
      ag.x += dy[4 * i] * xh[4 * i]; ag.y += dy[4 * i + 1] * xh[4 * i + 1];
      ag.z += dy[4 * i + 2] * xh[4 * i + 2]; ag.w += dy[4 * i + 3] * xh[4 * i + 3];
      ab.x += dy[4 * i]; ab.y += dy[4 * i + 1]; ab.z += dy[4 * i + 2]; ab.w += dy[4 * i + 3];
      *a_g = ag; *a_b = ab;
      // from here on dy holds dy * gamma
      dy[4 * i] *= g.x; dy[4 * i + 1] *= g.y; dy[4 * i + 2] *= g.z; dy[4 * i + 3] *= g.w;
#pragma unroll
      for (int e = 0; e < 4; ++e) { sa += dy[4 * i + e] * xh[4 * i + e]; sb += dy[4 * i + e]; }
    }
    sa = warp_sum(sa) * (1.0f / D);
    sb = warp_sum(sb) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int c = i * 128 + lane * 4;
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = rs * (dy[4 * i + e] - sb - xh[4 * i + e] * sa);
      if (g_in) { o[0] += gi[i].x; o[1] += gi[i].y; o[2] += gi[i].z; o[3] += gi[i].w; }
      if (g_out) *reinterpret_cast<float4*>(g_out + off + c) = make_float4(o[0], o[1], o[2], o[3]);
      float4* a_c = reinterpret_cast<float4*>(acc + 2 * D + c);
      float4 ac = *a_c;
      if (g_out_bf16) {
        uint2 ob;
        ob.x = pack_bf16(o[0], o[1]);
        ob.y = pack_bf16(o[2], o[3]);
        *reinterpret_cast<uint2*>(g_out_bf16 + off + c) = ob;
        // bias gradients are taken from the bf16 values the wgrad GEMM will also see
        const __nv_bfloat162 q0 = *reinterpret_cast<const __nv_bfloat162*>(&ob.x);
        const __nv_bfloat162 q1 = *reinterpret_cast<const __nv_bfloat162*>(&ob.y);
        ac.x += __low2float(q0); ac.y += __high2float(q0); ac.z += __low2float(q1); ac.w += __high2float(q1);
      } else {
        ac.x += o[0]; ac.y += o[1]; ac.z += o[2]; ac.w += o[3];
      }
      *a_c = ac;
    }
  }
  __syncthreads();
  // cross-warp reduction, then one atomic per column per CTA
  float* outs[3] = {dgamma, dbeta, colsum};
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    if (outs[which] == nullptr) continue;  // uniform
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kLnBwdWarps; ++w) s += acc_all[w * 3 * D + which * D + c];
      atomicAdd(outs[which] + c, s);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// column sums of a bf16 matrix (bias gradients): out[n] += sum_m in[m, n]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int M, int N, int ld,
              int rows_per_cta) {
  __shared__ float red[8][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 64 + lane * 2;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float a0 = 0.f, a1 = 0.f;
  if (c < N) {
    for (int r = r0 + warp; r < r1; r += 8) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(in + size_t(r) * ld + c);
      a0 += __low2float(v);
      a1 += __high2float(v);
    }
  }
  red[warp][lane * 2] = a0;
  red[warp][lane * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    const int cc = blockIdx.x * 64 + threadIdx.x;
    if (cc < N) atomicAdd(out + cc, s);
  }
}

// ------------------------------------------------------------------------------------------------
// im2col of the kept patches (optionally mixing two clips: x*lam + x[perm]*(1-lam), ex_audioset.py:173-177)
// ------------------------------------------------------------------------------------------------
template <bool F32OUT>
__global__ void __launch_bounds__(256)
im2col_kernel(const float* __restrict__ mel, void* __restrict__ A, const int* __restrict__ patch_f,
              const int* __restrict__ patch_t, int B, int ntok, int Fm, int Tm, int fstride, int tstride,
              const int* __restrict__ mix_perm, const float* __restrict__ mix_lam) {
  pdl_gate();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= B * ntok) return;
  const int b = gw / ntok, n = gw - b * ntok;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (n >= 2) {
    const int f0 = patch_f[n - 2] * fstride, t0 = patch_t[n - 2] * tstride;
    const int ky = lane >> 1, kx = (lane & 1) * 8;
    const float* src = mel + (size_t(b) * Fm + f0 + ky) * Tm + t0 + kx;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(src + i);
    if (mix_perm) {
      const float lam = mix_lam[b];
      const float* src2 = mel + (size_t(mix_perm[b]) * Fm + f0 + ky) * Tm + t0 + kx;
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = v[i] * lam + __ldg(src2 + i) * (1.0f - lam);
    }
  }
  if (F32OUT) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(A) + size_t(gw) * 256 + lane * 8);
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
    o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(A) + size_t(gw) * 256 + lane * 8) = o;
  }
}

// ------------------------------------------------------------------------------------------------
// token table: tab[n, c] (n < ntok)
//   n = 0: cls_token + new_pos_embed[0];  n = 1: dist_token + new_pos_embed[1]
//   n >= 2: conv_bias[c] + time_pos[c, toff + t(n)] + freq_pos[c, f(n)]
// ------------------------------------------------------------------------------------------------
__global__ void token_table_kernel(float* __restrict__ tab, const float* __restrict__ cls,
                                   const float* __restrict__ dist, const float* __restrict__ new_pos,
                                   const float* __restrict__ conv_bias, const float* __restrict__ time_pos,
                                   const float* __restrict__ freq_pos, const int* __restrict__ patch_f,
                                   const int* __restrict__ patch_t, int ntok, int Fg, int Tg, int toff,
                                   const int* __restrict__ toff_dev) {
  pdl_gate();
  const int n = blockIdx.x;
  if (toff_dev != nullptr) toff = *toff_dev;   // device-resident offset (CUDA-graph replays)
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float v;
    if (n == 0) v = cls[c] + new_pos[c];
    else if (n == 1) v = dist[c] + new_pos[D + c];
    else v = conv_bias[c] + time_pos[c * Tg + toff + patch_t[n - 2]] + freq_pos[c * Fg + patch_f[n - 2]];
    tab[size_t(n) * D + c] = v;
  }
}

// batch-sum of the block-0 input gradient per token, scattered to the table's sources
__global__ void __launch_bounds__(192)
token_table_bwd_kernel(const float* __restrict__ g0, float* __restrict__ dcls, float* __restrict__ ddist,
                       float* __restrict__ dnew_pos, float* __restrict__ dconv_bias, float* __restrict__ dtime,
                       float* __restrict__ dfreq, const int* __restrict__ patch_f, const int* __restrict__ patch_t,
                       int B, int ntok, int Fg, int Tg, int toff, const int* __restrict__ toff_dev) {
  pdl_gate();
  const int n = blockIdx.x;
  if (toff_dev != nullptr) toff = *toff_dev;
  const int c = threadIdx.x * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    const float4 v = *reinterpret_cast<const float4*>(g0 + (size_t(b) * ntok + n) * D + c);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  const float sv[4] = {s.x, s.y, s.z, s.w};
  if (n == 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { atomicAdd(dcls + c + e, sv[e]); atomicAdd(dnew_pos + c + e, sv[e]); }
  } else if (n == 1) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { atomicAdd(ddist + c + e, sv[e]); atomicAdd(dnew_pos + D + c + e, sv[e]); }
  } else {
    const int t = toff + patch_t[n - 2], f = patch_f[n - 2];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      atomicAdd(dconv_bias + c + e, sv[e]);
      atomicAdd(dtime + (c + e) * Tg + t, sv[e]);
      atomicAdd(dfreq + (c + e) * Fg + f, sv[e]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 [R, C] -> bf16 [R, C] and bf16 [C, R]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cast_transpose_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ outT,
                      int R, int C) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      v = in[size_t(r) * C + c];
      if (out) out[size_t(r) * C + c] = __float2bfloat16(v);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  if (outT) {
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i, r = r0 + tx;
      if (r < R && c < C) outT[size_t(c) * R + r] = __float2bfloat16(tile[tx][i]);
    }
  }
}

// fp32 -> bf16, 8 elements per thread (no transposed copy requested)
__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n8) {
  const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float4 a = reinterpret_cast<const float4*>(in)[2 * i];
  const float4 b = reinterpret_cast<const float4*>(in)[2 * i + 1];
  uint4 o;
  o.x = pack_bf16(a.x, a.y); o.y = pack_bf16(a.z, a.w); o.z = pack_bf16(b.x, b.y); o.w = pack_bf16(b.z, b.w);
  reinterpret_cast<uint4*>(out)[i] = o;
}

// fp32 -> bf16 for a whole list of matrices in one launch.  table[e] = {src, dst, n8, first_block}; block b works on
// 1024 8-element groups of the entry whose [first_block, next first_block) range contains it.
struct CastEntry {
  const float* src;
  __nv_bfloat16* dst;
  unsigned long long n8;
  unsigned int first_block, pad;
};
constexpr int kCastGroupsPerBlock = 1024;   // 256 threads x 4 groups of 8 elements
__global__ void __launch_bounds__(256)
cast_multi_kernel(const CastEntry* __restrict__ table, int n_entries) {
  pdl_gate();
  __shared__ int s_e;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_entries - 1;
    while (lo < hi) {                       // last entry with first_block <= blockIdx.x
      const int mid = (lo + hi + 1) >> 1;
      if (table[mid].first_block <= blockIdx.x) lo = mid; else hi = mid - 1;
    }
    s_e = lo;
  }
  __syncthreads();
  const CastEntry e = table[s_e];
  const size_t g0 = size_t(blockIdx.x - e.first_block) * kCastGroupsPerBlock;
  const float4* in = reinterpret_cast<const float4*>(e.src);
  uint4* out = reinterpret_cast<uint4*>(e.dst);
  float4 a[4], b[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const size_t i = g0 + u * 256 + threadIdx.x;
    if (i < e.n8) { a[u] = in[2 * i]; b[u] = in[2 * i + 1]; }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const size_t i = g0 + u * 256 + threadIdx.x;
    if (i < e.n8) {
      uint4 o;
      o.x = pack_bf16(a[u].x, a[u].y); o.y = pack_bf16(a[u].z, a[u].w);
      o.z = pack_bf16(b[u].x, b[u].y); o.w = pack_bf16(b[u].z, b[u].w);
      out[i] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// classifier head
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float v, float* scratch) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += scratch[w];
  return s;
}

struct HeadParams {
  const float* x;               // [B, ntok, D] residual stream before the last delta
  const __nv_bfloat16* delta;   // [B, ntok, D] last fc2 output or nullptr
  const float *norm_g, *norm_b; // final norm (eps 1e-6)
  const float *hln_g, *hln_b;   // head.0 LayerNorm (eps 1e-5)
  const float *W, *bias;        // head.1 Linear [C, D], [C]
  int B, ntok, C;
  float eps_norm, eps_head;
};

// one CTA (256 threads) per clip; thread owns columns tid, tid+256, tid+512
__global__ void __launch_bounds__(256)
head_fwd_kernel(const HeadParams p, float* __restrict__ logits, float* __restrict__ features,
                float* __restrict__ fl_out) {
  pdl_gate();
  __shared__ float scratch[8];
  __shared__ float s_fl[D];
  const int b = blockIdx.x, tid = threadIdx.x;
  float feat[3] = {0.f, 0.f, 0.f};
  for (int tok = 0; tok < 2; ++tok) {
    const size_t off = (size_t(b) * p.ntok + tok) * D;
    float v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = tid + 256 * i;
      v[i] = p.x[off + c] + (p.delta ? __bfloat162float(p.delta[off + c]) : 0.f);
    }
    const float mean = block_sum_256(v[0] + v[1] + v[2], scratch) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) q += (v[i] - mean) * (v[i] - mean);
    const float rstd = rsqrtf(block_sum_256(q, scratch) * (1.0f / D) + p.eps_norm);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = tid + 256 * i;
      feat[i] += 0.5f * ((v[i] - mean) * rstd * p.norm_g[c] + p.norm_b[c]);
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) features[size_t(b) * D + tid + 256 * i] = feat[i];
  const float mean = block_sum_256(feat[0] + feat[1] + feat[2], scratch) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) q += (feat[i] - mean) * (feat[i] - mean);
  const float rstd = rsqrtf(block_sum_256(q, scratch) * (1.0f / D) + p.eps_head);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = tid + 256 * i;
    const float f = (feat[i] - mean) * rstd * p.hln_g[c] + p.hln_b[c];
    s_fl[c] = f;
    if (fl_out) fl_out[size_t(b) * D + c] = f;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  // four classes per pass: 24 independent 16-byte loads in flight per lane (the weight rows come from L2)
  float fr[D / 128][4];
#pragma unroll
  for (int i = 0; i < D / 128; ++i) {
    const float4 f4 = *reinterpret_cast<const float4*>(s_fl + i * 128 + lane * 4);
    fr[i][0] = f4.x; fr[i][1] = f4.y; fr[i][2] = f4.z; fr[i][3] = f4.w;
  }
  for (int cls0 = warp * 4; cls0 < p.C; cls0 += 32) {
    float acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc[u] = 0.f;
      const int cls = min(cls0 + u, p.C - 1);
      const float* w = p.W + size_t(cls) * D;
#pragma unroll
      for (int i = 0; i < D / 128; ++i) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + i * 128 + lane * 4));
        acc[u] += wv.x * fr[i][0] + wv.y * fr[i][1] + wv.z * fr[i][2] + wv.w * fr[i][3];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float t = warp_sum(acc[u]);
      if (lane == 0 && cls0 + u < p.C) logits[size_t(b) * p.C + cls0 + u] = t + p.bias[cls0 + u];
    }
  }
}

// backward w.r.t. activations: writes g rows 0,1 of each clip (fp32 + bf16), LN parameter grads, colsum
__global__ void __launch_bounds__(256)
head_bwd_kernel(const HeadParams p, const float* __restrict__ dlogits, const float* __restrict__ dfeatures,
                float* __restrict__ g_out, __nv_bfloat16* __restrict__ g_out_bf16, float* __restrict__ d_norm_g,
                float* __restrict__ d_norm_b, float* __restrict__ d_hln_g, float* __restrict__ d_hln_b,
                float* __restrict__ colsum) {
  pdl_gate();
  __shared__ float scratch[8];
  __shared__ float s_dl[1024];
  const int b = blockIdx.x, tid = threadIdx.x;
  // ---- recompute forward
  float v[2][3], xh[2][3], rstd_tok[2];
  float feat[3] = {0.f, 0.f, 0.f};
  for (int tok = 0; tok < 2; ++tok) {
    const size_t off = (size_t(b) * p.ntok + tok) * D;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = tid + 256 * i;
      v[tok][i] = p.x[off + c] + (p.delta ? __bfloat162float(p.delta[off + c]) : 0.f);
    }
    const float mean = block_sum_256(v[tok][0] + v[tok][1] + v[tok][2], scratch) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) q += (v[tok][i] - mean) * (v[tok][i] - mean);
    const float rstd = rsqrtf(block_sum_256(q, scratch) * (1.0f / D) + p.eps_norm);
    rstd_tok[tok] = rstd;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = tid + 256 * i;
      xh[tok][i] = (v[tok][i] - mean) * rstd;
      feat[i] += 0.5f * (xh[tok][i] * p.norm_g[c] + p.norm_b[c]);
    }
  }
  const float fmean = block_sum_256(feat[0] + feat[1] + feat[2], scratch) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) q += (feat[i] - fmean) * (feat[i] - fmean);
  const float frstd = rsqrtf(block_sum_256(q, scratch) * (1.0f / D) + p.eps_head);
  // ---- d fl = dlogits . W
  for (int c = tid; c < p.C; c += 256) s_dl[c] = dlogits ? dlogits[size_t(b) * p.C + c] : 0.f;
  __syncthreads();
  // threads 0..191 own four consecutive columns each (one 16-byte load per class, eight classes in flight); the
  // result goes through shared memory back to the strided column ownership of the rest of the kernel
  __shared__ float s_dfl[D];
  if (tid < D / 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* w4 = reinterpret_cast<const float4*>(p.W) + tid;
#pragma unroll 8
    for (int cls = 0; cls < p.C; ++cls) {
      const float dl = s_dl[cls];
      const float4 w = __ldg(w4 + size_t(cls) * (D / 4));
      acc.x = fmaf(dl, w.x, acc.x); acc.y = fmaf(dl, w.y, acc.y);
      acc.z = fmaf(dl, w.z, acc.z); acc.w = fmaf(dl, w.w, acc.w);
    }
    *reinterpret_cast<float4*>(s_dfl + 4 * tid) = acc;
  }
  __syncthreads();
  float dfl[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) dfl[i] = s_dfl[tid + 256 * i];
  // ---- head LayerNorm backward
  float fh[3], dg[3];
  float sa = 0.f, sb = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = tid + 256 * i;
    fh[i] = (feat[i] - fmean) * frstd;
    atomicAdd(d_hln_g + c, dfl[i] * fh[i]);
    atomicAdd(d_hln_b + c, dfl[i]);
    dg[i] = dfl[i] * p.hln_g[c];
    sa += dg[i] * fh[i];
    sb += dg[i];
  }
  sa = block_sum_256(sa, scratch) * (1.0f / D);
  sb = block_sum_256(sb, scratch) * (1.0f / D);
  float dfeat[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    dfeat[i] = frstd * (dg[i] - sb - fh[i] * sa);
    if (dfeatures) dfeat[i] += dfeatures[size_t(b) * D + tid + 256 * i];
  }
  // ---- features = (y0 + y1)/2 ; y_tok = norm(v_tok)
  for (int tok = 0; tok < 2; ++tok) {
    float dy[3], dgn[3];
    float ta = 0.f, tb = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = tid + 256 * i;
      dy[i] = 0.5f * dfeat[i];
      atomicAdd(d_norm_g + c, dy[i] * xh[tok][i]);
      atomicAdd(d_norm_b + c, dy[i]);
      dgn[i] = dy[i] * p.norm_g[c];
      ta += dgn[i] * xh[tok][i];
      tb += dgn[i];
    }
    ta = block_sum_256(ta, scratch) * (1.0f / D);
    tb = block_sum_256(tb, scratch) * (1.0f / D);
    const size_t off = (size_t(b) * p.ntok + tok) * D;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = tid + 256 * i;
      const float g = rstd_tok[tok] * (dgn[i] - tb - xh[tok][i] * ta);
      g_out[off + c] = g;
      const __nv_bfloat16 gb = __float2bfloat16(g);
      g_out_bf16[off + c] = gb;
      if (colsum) atomicAdd(colsum + c, __bfloat162float(gb));
    }
  }
}

// dW[c, k] += sum_b dlogits[b, c] * fl[b, k];  db[c] += sum_b dlogits[b, c]
__global__ void __launch_bounds__(256)
head_wgrad_kernel(const float* __restrict__ dlogits, const float* __restrict__ fl, float* __restrict__ dW,
                  float* __restrict__ db, int B, int C) {
  pdl_gate();
  const int cls = blockIdx.x, tid = threadIdx.x;
  float acc[3] = {0.f, 0.f, 0.f};
  float sb = 0.f;
  for (int b = 0; b < B; ++b) {
    const float dl = dlogits[size_t(b) * C + cls];
    sb += dl;
#pragma unroll
    for (int i = 0; i < 3; ++i) acc[i] += dl * fl[size_t(b) * D + tid + 256 * i];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) dW[size_t(cls) * D + tid + 256 * i] += acc[i];
  if (tid == 0) db[cls] += sb;
}

}  // namespace pb

// ==================================================================================================
// C ABI
// ==================================================================================================
extern "C" {

int passt_ln_fwd(const float* x_in, const void* delta_bf16, float* x_out, void* h_bf16, float* mean, float* rstd,
                 const float* gamma, const float* beta, int M, int dim, float eps, void* stream) {
  using namespace pb;
  if (dim != D || M <= 0) return PB_ERR_BAD_ARG;
  const int threads = 256, rows_per_cta = threads / 32;
  PB_LAUNCH(ln_fwd_kernel, (M + rows_per_cta - 1) / rows_per_cta, threads, 0, (cudaStream_t)stream, x_in,
            (const __nv_bfloat16*)delta_bf16, x_out, (__nv_bfloat16*)h_bf16, mean, rstd, gamma, beta, M, eps);
  return 0;
}

int passt_ln_bwd(const void* dh_bf16, const float* x, const float* mean, const float* rstd, const float* gamma,
                 const float* g_in, float* g_out, void* g_out_bf16, float* dgamma, float* dbeta, float* colsum,
                 int M, int dim, void* stream) {
  using namespace pb;
  if (dim != D || M <= 0) return PB_ERR_BAD_ARG;
  int ctas = kNumSMs * 4;
  int rows_per_cta = (M + ctas - 1) / ctas;
  if (rows_per_cta < kLnBwdWarps) rows_per_cta = kLnBwdWarps;
  ctas = (M + rows_per_cta - 1) / rows_per_cta;
  PB_SET_SMEM_ONCE(kLnBwdSmem, ln_bwd_kernel);
  PB_LAUNCH(ln_bwd_kernel, ctas, kLnBwdWarps * 32, kLnBwdSmem, (cudaStream_t)stream, (const __nv_bfloat16*)dh_bf16, x,
            mean, rstd, gamma, g_in, g_out, (__nv_bfloat16*)g_out_bf16, dgamma, dbeta, colsum, M, rows_per_cta);
  return 0;
}

int passt_colsum_bf16(const void* in_bf16, float* out, int M, int N, int ld, void* stream) {
  using namespace pb;
  if (M <= 0 || N <= 0 || (N % 2) || (ld % 2)) return PB_ERR_BAD_ARG;
  const int col_blocks = (N + 63) / 64;
  int row_blocks = (kNumSMs * 4 + col_blocks - 1) / col_blocks;
  int rows_per_cta = (M + row_blocks - 1) / row_blocks;
  if (rows_per_cta < 8) rows_per_cta = 8;
  row_blocks = (M + rows_per_cta - 1) / rows_per_cta;
  colsum_kernel<<<dim3(col_blocks, row_blocks), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in_bf16, out,
                                                                                 M, N, ld, rows_per_cta);
  PB_LAUNCH_CHECK();
  return 0;
}

int passt_im2col(const float* mel, void* A_bf16, const int* patch_f, const int* patch_t, int B, int ntok, int Fm,
                 int Tm, int fstride, int tstride, const int* mix_perm, const float* mix_lam, void* stream) {
  using namespace pb;
  if (B <= 0 || ntok < 2) return PB_ERR_BAD_ARG;
  const long long warps = (long long)B * ntok;
  const int blocks = int((warps * 32 + 255) / 256);
  PB_LAUNCH(im2col_kernel<false>, blocks, 256, 0, (cudaStream_t)stream, mel, A_bf16, patch_f, patch_t, B, ntok, Fm, Tm,
            fstride, tstride, mix_perm, mix_lam);
  return 0;
}

// same gather with fp32 output rows [B*ntok, 256] (fp32-parity tier: the rows are split into bf16 hi/lo afterwards)
int passt_im2col_f32(const float* mel, float* A_f32, const int* patch_f, const int* patch_t, int B, int ntok, int Fm,
                     int Tm, int fstride, int tstride, const int* mix_perm, const float* mix_lam, void* stream) {
  using namespace pb;
  if (B <= 0 || ntok < 2) return PB_ERR_BAD_ARG;
  const long long warps = (long long)B * ntok;
  const int blocks = int((warps * 32 + 255) / 256);
  im2col_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(mel, A_f32, patch_f, patch_t, B, ntok, Fm, Tm, fstride,
                                                                tstride, mix_perm, mix_lam);
  PB_LAUNCH_CHECK();
  return 0;
}

int passt_token_table(float* tab, const float* cls, const float* dist, const float* new_pos,
                      const float* conv_bias, const float* time_pos, const float* freq_pos, const int* patch_f,
                      const int* patch_t, int ntok, int Fg, int Tg, int toff, const int* toff_dev, void* stream) {
  using namespace pb;
  if (ntok < 2) return PB_ERR_BAD_ARG;
  PB_LAUNCH(token_table_kernel, ntok, 256, 0, (cudaStream_t)stream, tab, cls, dist, new_pos, conv_bias, time_pos,
            freq_pos, patch_f, patch_t, ntok, Fg, Tg, toff, toff_dev);
  return 0;
}

int passt_token_table_bwd(const float* g0, float* dcls, float* ddist, float* dnew_pos, float* dconv_bias,
                          float* dtime, float* dfreq, const int* patch_f, const int* patch_t, int B, int ntok,
                          int Fg, int Tg, int toff, const int* toff_dev, void* stream) {
  using namespace pb;
  if (ntok < 2 || B <= 0) return PB_ERR_BAD_ARG;
  PB_LAUNCH(token_table_bwd_kernel, ntok, 192, 0, (cudaStream_t)stream, g0, dcls, ddist, dnew_pos, dconv_bias, dtime,
            dfreq, patch_f, patch_t, B, ntok, Fg, Tg, toff, toff_dev);
  return 0;
}

int passt_cast_transpose(const float* in, void* out_bf16, void* outT_bf16, int R, int C, void* stream) {
  using namespace pb;
  if (R <= 0 || C <= 0) return PB_ERR_BAD_ARG;
  if (outT_bf16 == nullptr && out_bf16 != nullptr && (size_t(R) * C) % 8 == 0) {
    const size_t n8 = size_t(R) * C / 8;
    cast_bf16_kernel<<<unsigned((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out_bf16, n8);
    PB_LAUNCH_CHECK();
    return 0;
  }
  cast_transpose_kernel<<<dim3((C + 31) / 32, (R + 31) / 32), 256, 0, (cudaStream_t)stream>>>(
      in, (__nv_bfloat16*)out_bf16, (__nv_bfloat16*)outT_bf16, R, C);
  PB_LAUNCH_CHECK();
  return 0;
}

// table: device array of n_entries records {const float* src; bf16* dst; uint64 n8; uint32 first_block; uint32 pad}
// (32 bytes each; n8 = elements / 8; first_block = running sum of ceil(n8 / 1024)); total_blocks = that sum's end.
int passt_cast_multi(const void* table, int n_entries, int total_blocks, void* stream) {
  using namespace pb;
  static_assert(sizeof(CastEntry) == 32, "CastEntry layout is part of the C ABI");
  if (table == nullptr || n_entries <= 0 || total_blocks <= 0) return PB_ERR_BAD_ARG;
  PB_LAUNCH(cast_multi_kernel, total_blocks, 256, 0, (cudaStream_t)stream, reinterpret_cast<const CastEntry*>(table),
            n_entries);
  return 0;
}

int passt_head_fwd(const float* x, const void* delta_bf16, const float* norm_g, const float* norm_b,
                   const float* hln_g, const float* hln_b, const float* W, const float* bias, float* logits,
                   float* features, float* fl, int B, int ntok, int C, void* stream) {
  using namespace pb;
  if (B <= 0 || C <= 0 || C > 1024) return PB_ERR_BAD_ARG;
  HeadParams p{x, (const __nv_bfloat16*)delta_bf16, norm_g, norm_b, hln_g, hln_b, W, bias, B, ntok, C, 1e-6f, 1e-5f};
  PB_LAUNCH(head_fwd_kernel, B, 256, 0, (cudaStream_t)stream, p, logits, features, fl);
  return 0;
}

int passt_head_bwd(const float* x, const void* delta_bf16, const float* norm_g, const float* norm_b,
                   const float* hln_g, const float* hln_b, const float* W, const float* dlogits,
                   const float* dfeatures, const float* fl, float* g_out, void* g_out_bf16, float* d_norm_g,
                   float* d_norm_b, float* d_hln_g, float* d_hln_b, float* dW, float* dbias, float* colsum, int B,
                   int ntok, int C, void* stream) {
  using namespace pb;
  if (B <= 0 || C <= 0 || C > 1024) return PB_ERR_BAD_ARG;
  HeadParams p{x, (const __nv_bfloat16*)delta_bf16, norm_g, norm_b, hln_g, hln_b, W, nullptr, B, ntok, C, 1e-6f, 1e-5f};
  PB_LAUNCH(head_bwd_kernel, B, 256, 0, (cudaStream_t)stream, p, dlogits, dfeatures, g_out, (__nv_bfloat16*)g_out_bf16,
            d_norm_g, d_norm_b, d_hln_g, d_hln_b, colsum);
  if (dlogits) {
    PB_LAUNCH(head_wgrad_kernel, C, 256, 0, (cudaStream_t)stream, dlogits, fl, dW, dbias, B, C);
  }
  return 0;
}

}  // extern "C"
