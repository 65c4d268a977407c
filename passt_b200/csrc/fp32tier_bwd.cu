// passt_b200 — backward pass of the fp32-parity tier (north_star: gradients within 1e-3 of the fp32 reference).
//
// Same recipe as the forward (fp32tier.cu): every contraction runs on the bf16 tcgen05 tensor cores over hi/lo-split
// operands with a 3x longer contraction and fp32 accumulation / output; everything else is fp32 on the CUDA cores.
//   dgrad   dX = dY W        : A' = [dY_hi | dY_hi | dY_lo] (columns), B' = [W_hi ; W_lo ; W_hi] (rows)   -> passt_gemm_bf16 mode 2|KN
//   wgrad   dW += dY^T X     : A' = [dY_hi ; dY_hi ; dY_lo] (rows),    B' = [X_hi ; X_lo ; X_hi] (rows)   -> passt_gemm_bf16 mode 4
// This file holds the fp32 glue kernels: row-stacked splits, column sums (bias gradients), LayerNorm forward
// recomputation / backward, GELU backward and the attention backward (models/passt.py:271-380 autograd).
#include "common.cuh"

namespace pb {

__device__ __forceinline__ void split_bf16_b(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16(x);
  lo = __float2bfloat16(x - __bfloat162float(hi));
}

// fp32 [R, C] (row stride ld_in) -> bf16 [3R, C]: rows [0,R) / [R,2R) / [2R,3R) hold hi / (hi or lo) / (lo or hi)
// pattern 0 = [hi; hi; lo] (the "A" side of a split product), pattern 1 = [hi; lo; hi] (the "B" side)
__global__ void __launch_bounds__(256)
split3_rows_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long R, int C, int ld_in,
                   int pattern) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= R * C) return;
  const long long r = i / C;
  const int c = int(i - r * C);
  __nv_bfloat16 hi, lo;
  split_bf16_b(in[r * ld_in + c], hi, lo);
  out[r * C + c] = hi;
  out[(R + r) * C + c] = pattern == 0 ? hi : lo;
  out[(2 * R + r) * C + c] = pattern == 0 ? lo : hi;
}

// column-stacked split operand saved by the forward ([R, 3C] = [hi | hi | lo]) -> row-stacked [3R, C] in pattern 0 or 1
__global__ void __launch_bounds__(256)
restack3_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long R, int C, int pattern) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= R * C) return;
  const long long r = i / C;
  const int c = int(i - r * C);
  const __nv_bfloat16 hi = in[r * 3 * C + c], lo = in[r * 3 * C + 2 * C + c];
  out[r * C + c] = hi;
  out[(R + r) * C + c] = pattern == 0 ? hi : lo;
  out[(2 * R + r) * C + c] = pattern == 0 ? lo : hi;
}

// out[c] += sum_r in[r, c]   (fp32; 32 columns per CTA column block, rows strided over blockIdx.y)
__global__ void __launch_bounds__(256)
colsum_f32_kernel(const float* __restrict__ in, float* __restrict__ out, long long R, int C, int rows_per_cta) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < R ? r0 + rows_per_cta : R;
  float a = 0.f;
  if (c < C)
    for (long long r = r0 + w; r < r1; r += 8) a += in[r * C + c];
  red[w][lane] = a;
  __syncthreads();
  if (w == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][lane];
    atomicAdd(out + c, s);
  }
}

// h = LayerNorm(x) * gamma + beta in fp32 (recomputed for the weight-gradient products)   [dim 768, warp per row]
__global__ void __launch_bounds__(256)
ln_apply_f32_kernel(const float* __restrict__ x, float* __restrict__ h, const float* __restrict__ gamma,
                    const float* __restrict__ beta, int M, float eps) {
  constexpr int Dm = 768;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const size_t off = size_t(row) * Dm;
  float v[24];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { v[i] = x[off + i * 32 + lane]; s += v[i]; }
  const float mean = warp_sum(s) * (1.0f / Dm);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / Dm) + eps);
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const int c = i * 32 + lane;
    h[off + c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
  }
}

// g_out = g_in + dLN(dh; x, gamma);  dgamma += sum dh * xhat;  dbeta += sum dh   (all fp32; warp per row, CTA-level
// column partials in shared memory, one atomic per column and CTA)
__global__ void __launch_bounds__(256)
ln_bwd_f32_kernel(const float* __restrict__ dh, const float* __restrict__ x, const float* __restrict__ gamma,
                  const float* g_in, float* g_out, float* __restrict__ dgamma, float* __restrict__ dbeta, int M,
                  float eps, int rows_per_cta) {
  constexpr int Dm = 768;
  __shared__ float s_dg[Dm], s_db[Dm];
  for (int i = threadIdx.x; i < Dm; i += 256) { s_dg[i] = 0.f; s_db[i] = 0.f; }
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float adg[24], adb[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) { adg[i] = 0.f; adb[i] = 0.f; }
  for (int row = r0 + w; row < r1; row += 8) {
    const size_t off = size_t(row) * Dm;
    float v[24], dy[24];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) { v[i] = x[off + i * 32 + lane]; s += v[i]; }
    const float mean = warp_sum(s) * (1.0f / Dm);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) { v[i] -= mean; q += v[i] * v[i]; }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / Dm) + eps);
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      const int c = i * 32 + lane;
      v[i] *= rstd;                                   // xhat
      const float d = dh[off + c];
      adg[i] += d * v[i];
      adb[i] += d;
      dy[i] = d * gamma[c];
      sa += dy[i] * v[i];
      sb += dy[i];
    }
    sa = warp_sum(sa) * (1.0f / Dm);
    sb = warp_sum(sb) * (1.0f / Dm);
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      const int c = i * 32 + lane;
      float o = rstd * (dy[i] - sb - v[i] * sa);
      if (g_in) o += g_in[off + c];
      g_out[off + c] = o;
    }
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    atomicAdd(&s_dg[i * 32 + lane], adg[i]);
    atomicAdd(&s_db[i * 32 + lane], adb[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Dm; i += 256) {
    atomicAdd(dgamma + i, s_dg[i]);
    atomicAdd(dbeta + i, s_db[i]);
  }
}

// dpre = dact * gelu'(pre)   (exact-erf GELU, models/passt.py:280)
__global__ void __launch_bounds__(256)
gelu_bwd_f32_kernel(const float* __restrict__ dact, const float* __restrict__ pre, float* __restrict__ dpre, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float x = pre[i];
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
  dpre[i] = dact[i] * (cdf + x * pdf);
}

// ---- attention backward in fp32 (head_dim 64), two simple passes over a recomputed probability matrix:
//   pass 1 (row pass, CTA = 16 query rows of one (clip, head), 256 threads = 16 threads per row):
//       m_i, l_i of softmax row i;  D_i = sum_j P_ij dP_ij  with dP_ij = <dO_i, V_j>;  dQ_i = scale * sum_j dS_ij K_j
//   pass 2 (column pass, CTA = 16 key rows):  dV_j = sum_i P_ij dO_i;  dK_j = scale * sum_i dS_ij Q_i
// with P_ij = exp(scale q_i.k_j - m_i) / l_i and dS_ij = P_ij (dP_ij - D_i).  O(N^2 d) work per head twice; this tier
// is for parity runs, not throughput.
constexpr int kAB = 16;     // rows per CTA
struct AttnBwdF32Params {
  const float* qkv;   // [B, N, 3C]
  const float* dO;    // [B, N, C]
  float* dqkv;        // [B, N, 3C]
  float* m;           // [B, H, N] row maxima (scaled scores)
  float* l;           // [B, H, N] row sums
  float* Dsum;        // [B, H, N]
  int N, H;
  float scale;
};

__global__ void __launch_bounds__(256)
attn_bwd_f32_rows_kernel(const AttnBwdF32Params p) {
  __shared__ float sq[kAB][64], sdo[kAB][64], sk[64][65], sv[64][65];
  __shared__ float s_m[kAB], s_l[kAB], s_D[kAB];
  const int C = p.H * 64;
  const int i0 = blockIdx.x * kAB, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, ri = tid >> 4, part = tid & 15;    // 16 threads per query row
  const float* base = p.qkv + size_t(b) * p.N * 3 * C;
  const float* dob = p.dO + size_t(b) * p.N * C;
  for (int t = tid; t < kAB * 64; t += 256) {
    const int r = t >> 6, d = t & 63;
    const bool ok = i0 + r < p.N;
    sq[r][d] = ok ? base[size_t(i0 + r) * 3 * C + h * 64 + d] : 0.f;
    sdo[r][d] = ok ? dob[size_t(i0 + r) * C + h * 64 + d] : 0.f;
  }
  __syncthreads();
  // ---- sweep 1: row max and sum
  float mx = -INFINITY;
  for (int j0 = 0; j0 < p.N; j0 += 64) {
    for (int t = tid; t < 64 * 64; t += 256) {
      const int r = t >> 6, d = t & 63;
      sk[r][d] = (j0 + r < p.N) ? base[size_t(j0 + r) * 3 * C + C + h * 64 + d] : 0.f;
    }
    __syncthreads();
    for (int jj = part; jj < 64; jj += 16) {
      if (j0 + jj >= p.N) break;
      float s = 0.f;
#pragma unroll 16
      for (int d = 0; d < 64; ++d) s = fmaf(sq[ri][d], sk[jj][d], s);
      mx = fmaxf(mx, s * p.scale);
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float lsum = 0.f, dacc = 0.f;
  float dq[4] = {0.f, 0.f, 0.f, 0.f};      // this thread's 4 of the 64 dQ columns: d = part*4 .. part*4+3
  // ---- sweep 2: l, D (needs P and dP), then sweep 3 for dQ needs D: do l and the unnormalised D first
  for (int j0 = 0; j0 < p.N; j0 += 64) {
    for (int t = tid; t < 64 * 64; t += 256) {
      const int r = t >> 6, d = t & 63;
      const bool ok = j0 + r < p.N;
      const size_t off = size_t(j0 + r) * 3 * C + h * 64 + d;
      sk[r][d] = ok ? base[off + C] : 0.f;
      sv[r][d] = ok ? base[off + 2 * C] : 0.f;
    }
    __syncthreads();
    for (int jj = part; jj < 64; jj += 16) {
      if (j0 + jj >= p.N) break;
      float s = 0.f, dp = 0.f;
#pragma unroll 16
      for (int d = 0; d < 64; ++d) { s = fmaf(sq[ri][d], sk[jj][d], s); dp = fmaf(sdo[ri][d], sv[jj][d], dp); }
      const float e = expf(s * p.scale - mx);
      lsum += e;
      dacc += e * dp;
    }
    __syncthreads();
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    dacc += __shfl_xor_sync(0xffffffffu, dacc, o);
  }
  const float Di = dacc / lsum;
  if (part == 0) { s_m[ri] = mx; s_l[ri] = lsum; s_D[ri] = Di; }
  __syncthreads();
  if (tid < kAB && i0 + tid < p.N) {
    const size_t o = (size_t(b) * p.H + h) * p.N + i0 + tid;
    p.m[o] = s_m[tid]; p.l[o] = s_l[tid]; p.Dsum[o] = s_D[tid];
  }
  // ---- sweep 3: dQ_i = scale * sum_j dS_ij K_j ; each thread accumulates 4 columns over ALL keys of the tile, so the
  //      16 threads of a row first share their dS values through shared memory
  __shared__ float s_ds[kAB][64];
  for (int j0 = 0; j0 < p.N; j0 += 64) {
    for (int t = tid; t < 64 * 64; t += 256) {
      const int r = t >> 6, d = t & 63;
      const bool ok = j0 + r < p.N;
      const size_t off = size_t(j0 + r) * 3 * C + h * 64 + d;
      sk[r][d] = ok ? base[off + C] : 0.f;
      sv[r][d] = ok ? base[off + 2 * C] : 0.f;
    }
    __syncthreads();
    for (int jj = part; jj < 64; jj += 16) {
      float ds = 0.f;
      if (j0 + jj < p.N) {
        float s = 0.f, dp = 0.f;
#pragma unroll 16
        for (int d = 0; d < 64; ++d) { s = fmaf(sq[ri][d], sk[jj][d], s); dp = fmaf(sdo[ri][d], sv[jj][d], dp); }
        const float pij = expf(s * p.scale - mx) / lsum;
        ds = pij * (dp - Di);
      }
      s_ds[ri][jj] = ds;
    }
    __syncthreads();
    for (int jj = 0; jj < 64; ++jj) {
      const float ds = s_ds[ri][jj];
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[e] = fmaf(ds, sk[jj][part * 4 + e], dq[e]);
    }
    __syncthreads();
  }
  if (i0 + ri < p.N) {
    float* o = p.dqkv + (size_t(b) * p.N + i0 + ri) * 3 * C + h * 64 + part * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = dq[e] * p.scale;
  }
}

__global__ void __launch_bounds__(256)
attn_bwd_f32_cols_kernel(const AttnBwdF32Params p) {
  constexpr int kQT = 32;     // queries per sweep step
  __shared__ float sk[kAB][64], sv[kAB][64], sqt[kQT][65], sdot[kQT][65];
  __shared__ float s_m[kQT], s_l[kQT], s_D[kQT];
  __shared__ float s_p[kAB][kQT], s_ds[kAB][kQT];
  const int C = p.H * 64;
  const int j0 = blockIdx.x * kAB, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, rj = tid >> 4, part = tid & 15;    // 16 threads per key row
  const float* base = p.qkv + size_t(b) * p.N * 3 * C;
  const float* dob = p.dO + size_t(b) * p.N * C;
  for (int t = tid; t < kAB * 64; t += 256) {
    const int r = t >> 6, d = t & 63;
    const bool ok = j0 + r < p.N;
    const size_t off = size_t(j0 + r) * 3 * C + h * 64 + d;
    sk[r][d] = ok ? base[off + C] : 0.f;
    sv[r][d] = ok ? base[off + 2 * C] : 0.f;
  }
  float dk[4] = {0.f, 0.f, 0.f, 0.f}, dv[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i0 = 0; i0 < p.N; i0 += kQT) {
    __syncthreads();
    for (int t = tid; t < kQT * 64; t += 256) {
      const int r = t >> 6, d = t & 63;
      const bool ok = i0 + r < p.N;
      sqt[r][d] = ok ? base[size_t(i0 + r) * 3 * C + h * 64 + d] : 0.f;
      sdot[r][d] = ok ? dob[size_t(i0 + r) * C + h * 64 + d] : 0.f;
    }
    if (tid < kQT) {
      const bool ok = i0 + tid < p.N;
      const size_t o = (size_t(b) * p.H + h) * p.N + i0 + tid;
      s_m[tid] = ok ? p.m[o] : 0.f; s_l[tid] = ok ? p.l[o] : 1.f; s_D[tid] = ok ? p.Dsum[o] : 0.f;
    }
    __syncthreads();
    for (int ii = part; ii < kQT; ii += 16) {
      float pij = 0.f, ds = 0.f;
      if (i0 + ii < p.N && j0 + rj < p.N) {
        float s = 0.f, dp = 0.f;
#pragma unroll 16
        for (int d = 0; d < 64; ++d) { s = fmaf(sqt[ii][d], sk[rj][d], s); dp = fmaf(sdot[ii][d], sv[rj][d], dp); }
        pij = expf(s * p.scale - s_m[ii]) / s_l[ii];
        ds = pij * (dp - s_D[ii]);
      }
      s_p[rj][ii] = pij;
      s_ds[rj][ii] = ds;
    }
    __syncthreads();
    for (int ii = 0; ii < kQT; ++ii) {
      const float pij = s_p[rj][ii], ds = s_ds[rj][ii];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dv[e] = fmaf(pij, sdot[ii][part * 4 + e], dv[e]);
        dk[e] = fmaf(ds, sqt[ii][part * 4 + e], dk[e]);
      }
    }
  }
  if (j0 + rj < p.N) {
    float* o = p.dqkv + (size_t(b) * p.N + j0 + rj) * 3 * C + h * 64 + part * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) { o[C + e] = dk[e] * p.scale; o[2 * C + e] = dv[e]; }
  }
}

}  // namespace pb

extern "C" {

// fp32 [R, C] (row stride ld_in) -> bf16 [3R, C]; pattern 0 = [hi; hi; lo], 1 = [hi; lo; hi]  (contraction along rows)
int passt_split3_rows_bf16(const float* in, void* out_bf16, long long R, int C, int ld_in, int pattern, void* stream) {
  using namespace pb;
  if (!in || !out_bf16 || R <= 0 || C <= 0 || ld_in < C || (pattern != 0 && pattern != 1)) return PB_ERR_BAD_ARG;
  const long long n = R * C;
  split3_rows_kernel<<<unsigned((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out_bf16, R, C,
                                                                                  ld_in, pattern);
  PB_LAUNCH_CHECK();
  return 0;
}

// bf16 [R, 3C] = [hi | hi | lo] (as written by passt_split3_bf16 pattern 0 / the fp32-tier forward) -> bf16 [3R, C] rows
int passt_restack3_bf16(const void* in_bf16, void* out_bf16, long long R, int C, int pattern, void* stream) {
  using namespace pb;
  if (!in_bf16 || !out_bf16 || R <= 0 || C <= 0 || (pattern != 0 && pattern != 1)) return PB_ERR_BAD_ARG;
  const long long n = R * C;
  restack3_kernel<<<unsigned((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in_bf16,
                                                                               (__nv_bfloat16*)out_bf16, R, C, pattern);
  PB_LAUNCH_CHECK();
  return 0;
}

// out[c] += sum_r in[r, c]  (fp32 bias gradients of the fp32 tier)
int passt_colsum_f32(const float* in, float* out, long long R, int C, void* stream) {
  using namespace pb;
  if (!in || !out || R <= 0 || C <= 0) return PB_ERR_BAD_ARG;
  const int col_blocks = (C + 31) / 32;
  long long row_blocks = (long long)(kNumSMs * 4 + col_blocks - 1) / col_blocks;
  long long rows_per_cta = (R + row_blocks - 1) / row_blocks;
  if (rows_per_cta < 8) rows_per_cta = 8;
  row_blocks = (R + rows_per_cta - 1) / rows_per_cta;
  colsum_f32_kernel<<<dim3(col_blocks, unsigned(row_blocks)), 256, 0, (cudaStream_t)stream>>>(in, out, R, C,
                                                                                             int(rows_per_cta));
  PB_LAUNCH_CHECK();
  return 0;
}

int passt_ln_apply_f32(const float* x, float* h, const float* gamma, const float* beta, int M, int dim, float eps,
                       void* stream) {
  using namespace pb;
  if (!x || !h || !gamma || !beta || M <= 0 || dim != 768) return PB_ERR_BAD_ARG;
  ln_apply_f32_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, h, gamma, beta, M, eps);
  PB_LAUNCH_CHECK();
  return 0;
}

// g_out = g_in (nullable) + dLN(dh; x, gamma); dgamma / dbeta += ...   (all fp32, dim 768)
int passt_ln_bwd_f32(const float* dh, const float* x, const float* gamma, const float* g_in, float* g_out, float* dgamma,
                     float* dbeta, int M, int dim, float eps, void* stream) {
  using namespace pb;
  if (!dh || !x || !gamma || !g_out || !dgamma || !dbeta || M <= 0 || dim != 768) return PB_ERR_BAD_ARG;
  int ctas = kNumSMs * 2;
  int rows_per_cta = (M + ctas - 1) / ctas;
  if (rows_per_cta < 8) rows_per_cta = 8;
  ctas = (M + rows_per_cta - 1) / rows_per_cta;
  ln_bwd_f32_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(dh, x, gamma, g_in, g_out, dgamma, dbeta, M, eps, rows_per_cta);
  PB_LAUNCH_CHECK();
  return 0;
}

int passt_gelu_bwd_f32(const float* dact, const float* pre, float* dpre, long long n, void* stream) {
  using namespace pb;
  if (!dact || !pre || !dpre || n <= 0) return PB_ERR_BAD_ARG;
  gelu_bwd_f32_kernel<<<unsigned((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dact, pre, dpre, n);
  PB_LAUNCH_CHECK();
  return 0;
}

// qkv f32 [B,N,3C], dO f32 [B,N,C] -> dqkv f32 [B,N,3C]; workspace: 3 * B*H*N floats
int passt_attn_bwd_f32(const float* qkv, const float* dO, float* dqkv, float* workspace, int B, int N, int H, float scale,
                       void* stream) {
  using namespace pb;
  if (!qkv || !dO || !dqkv || !workspace || B <= 0 || N <= 0 || H <= 0) return PB_ERR_BAD_ARG;
  AttnBwdF32Params p;
  p.qkv = qkv; p.dO = dO; p.dqkv = dqkv; p.N = N; p.H = H; p.scale = scale;
  p.m = workspace; p.l = workspace + size_t(B) * H * N; p.Dsum = workspace + 2 * size_t(B) * H * N;
  dim3 grid((N + kAB - 1) / kAB, H, B);
  attn_bwd_f32_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  PB_LAUNCH_CHECK();
  attn_bwd_f32_cols_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  PB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
