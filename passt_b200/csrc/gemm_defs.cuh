// passt_b200 — definitions shared by the 1-CTA and 2-CTA tcgen05 GEMM kernels.
#pragma once
#include "common.cuh"

namespace pb {

enum GemmMode : int {
  kBiasBf16 = 0,      // C = bf16(acc + bias)
  kBiasGeluBf16 = 1,  // C = bf16(gelu'(acc + bias)), C2 = bf16(gelu(acc + bias))
  kRowTabF32 = 2,     // C = fp32(acc + tab[row % period, col] (+ bias[col]))  -- period >= M: a full residual tensor
  kGeluGradBf16 = 3,  // C = bf16(acc * aux[row, col])   (aux = gelu'(pre) saved by mode 1)
  kWgradF32 = 4,      // C += fp32(acc)   (MN-major operands, split-K, TMA reduce-add)
  kRowDotBf16 = 5,    // C = bf16(acc);  dsum[(row / period) , col / 64, row % period] += sum_{64-col group} C * aux
                      //   (attention backward's D = rowsum(dO o O) folded into the proj-dgrad GEMM; needs kBRowMajorKN)
  kBRowMajorKN = 16,  // flag for modes 0 / 3: B is given as [K, N] row-major (the nn.Linear weight itself for dgrad)
};

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kStages = 3;
constexpr int kEpiWarps = 12;                      // 3 per TMEM lane quadrant; column chunks are dealt round-robin
constexpr int kEpiSlots = kEpiWarps / 4;
constexpr int kGemmThreads = 64 + kEpiWarps * 32;  // warp 0 TMA, warp 1 MMA, warps 2..13 epilogue

struct GemmParams {
  int M, N, K;             // D is [M,N]; K = contraction length
  int m_tiles, n_tiles;    // tile grid
  int k_blocks;            // total BK blocks along K
  int splits;              // split-K factor (1 for TN modes)
  const float* bias;       // [N] or nullptr.  mode 3: OUTPUT (float*), += column sums of C (bias gradient) if non-null
                           // mode 5: OUTPUT (float*) dsum [M / period, N / 64, npad] (npad = 128 * ceil(period / 128))
  const void* aux;         // mode 2: float tab[period, N]; mode 3: bf16 gelu'(pre) [M, ld_aux]; mode 5: bf16 O [M, ld_aux]
  int aux_period;          // mode 2: table rows; mode 5: tokens per clip
  int ld_aux;              // elements
  // UMMA smem-descriptor strides; exposed so the bring-up test can probe alternatives without a rebuild
  uint32_t lbo_a, sbo_a, kstep_a;  // bytes
  uint32_t lbo_b, sbo_b, kstep_b;  // bytes
};

__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// exact-erf GELU and its derivative from one rcp + one ex2 (Abramowitz-Stegun 7.1.26, |erf error| <= 1.5e-7):
//   gelu(x) = x * Phi(x),  gelu'(x) = Phi(x) + x * phi(x),  Phi = 0.5 (1 + erf(x / sqrt 2)),  phi = e^{-x^2/2}/sqrt(2 pi)
__device__ __forceinline__ void gelu_and_grad(float x, float& g, float& dg) {
  const float u = x * 0.70710678118654752f;
  const float t = fast_rcp(fmaf(0.3275911f, fabsf(u), 1.0f));
  const float e = fast_ex2(u * u * -1.4426950408889634f);            // exp(-u^2) = exp(-x^2/2)
  float q = fmaf(1.061405429f, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  const float erf_abs = fmaf(-q * t, e, 1.0f);
  const float cdf = fmaf(0.5f, copysignf(erf_abs, u), 0.5f);
  g = x * cdf;
  dg = fmaf(x * e, 0.3989422804014327f, cdf);
}


// ---- packed fp32 pairs (sm_100a FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot) ----
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rc, {%6, %7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0, %1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmul.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0, %1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nadd.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0, %1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// gelu_and_grad on two values at once.  Same Abramowitz-Stegun 7.1.26 evaluation, rearranged so that everything
// except the two rcp and two ex2 runs as packed FFMA2/FMUL2 (11 issue slots per element instead of 20):
//   k = sqrt(log2(e)/2),  y = k|x|  (so exp(-x^2/2) = 2^(-y^2)),  t = 1/(1 + (p/(k sqrt 2)) y),
//   w = (Phi(|x|) - 1/2)/k = 1/(2k) - (1/(2k)) (a1 t + ... + a5 t^5) 2^(-y^2),   Phi(x) = 1/2 + copysign(k, x) w.
__device__ __forceinline__ void gelu_and_grad2(float2 x, float2& g, float2& dg) {
  constexpr float k = 0.84932180028801904f;                       // sqrt(0.5 * log2(e))
  constexpr float pk = 0.3275911f * 0.70710678118654752f / k;
  constexpr float hk = 0.5f / k;
  const float2 ks = make_float2(__uint_as_float((__float_as_uint(x.x) & 0x80000000u) | __float_as_uint(k)),
                                __uint_as_float((__float_as_uint(x.y) & 0x80000000u) | __float_as_uint(k)));
  const float2 y = mul2(x, ks);
  const float2 d = fma2(y, make_float2(pk, pk), make_float2(1.f, 1.f));
  const float2 t = make_float2(fast_rcp(d.x), fast_rcp(d.y));
  const float2 yy = mul2(y, y);
  const float2 e = make_float2(fast_ex2(-yy.x), fast_ex2(-yy.y));
  float2 q = fma2(make_float2(-1.061405429f * hk, -1.061405429f * hk), t, make_float2(1.453152027f * hk, 1.453152027f * hk));
  q = fma2(q, t, make_float2(-1.421413741f * hk, -1.421413741f * hk));
  q = fma2(q, t, make_float2(0.284496736f * hk, 0.284496736f * hk));
  q = fma2(q, t, make_float2(-0.254829592f * hk, -0.254829592f * hk));
  const float2 w = fma2(mul2(q, t), e, make_float2(hk, hk));
  const float2 cdf = fma2(ks, w, make_float2(0.5f, 0.5f));
  g = mul2(x, cdf);
  dg = fma2(mul2(x, e), make_float2(0.3989422804014327f, 0.3989422804014327f), cdf);
}

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

struct DescOverride {
  int active = 0;
  uint32_t v[6];
};
extern DescOverride g_desc_override;

// 2-CTA (cta_group::2) variant, gemm2.cu.  Returns PB_ERR_BAD_ARG when the shape is not supported.
int launch_gemm2(int mode, const void* A, const void* B, void* C, void* C2, const float* bias, const void* aux, int M,
                 int N, int K, int lda, int ldb, int ldc, int aux_period, int ld_aux, int splits, cudaStream_t stream);

}  // namespace pb
