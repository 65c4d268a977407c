// passt_b200 — AdamW over a list of parameter tensors in one launch, with the bf16 GEMM-operand copy of each weight
// matrix refreshed in the same pass (sm_100a; HBM-bound: 28 B read+written per element, +2 B where a bf16 copy exists).
//
// Replaces, on the training hot path, the optimizer the reference builds in get_optimizer (ex_audioset.py:104-109:
// torch.optim.AdamW(params, lr, weight_decay)) plus the per-op autocast casts of the weights.  Update rule and operation
// order are those of torch's AdamW (decoupled weight decay, bias-corrected moments, no amsgrad):
//     p <- p - lr*wd*p;  m <- m + (1-b1)(g-m);  v <- b2*v + (1-b2) g^2;
//     p <- p - (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
// The step count t lives on the device (CUDA-graph replays advance it); lr is read from device memory every step.
#include "common.cuh"

namespace pb {

struct AdamEntry {                 // 64 bytes, part of the C ABI (built by passt_b200/optim.py)
  float* p;
  const float* g;
  float* m;
  float* v;
  __nv_bfloat16* w16;              // bf16 copy to refresh, or nullptr
  unsigned long long n;            // elements
  unsigned int first_block;        // running sum of ceil(n / 4096)
  unsigned int vec;                // 1: n % 4 == 0 and p, g, m, v 16-byte aligned (w16 8-byte aligned)
  unsigned long long reserved;
};
static_assert(sizeof(AdamEntry) == 64, "AdamEntry layout is part of the C ABI");

constexpr int kAdamElemsPerBlock = 4096;   // 256 threads x 4 groups x 4 elements

// hyper: [0] lr  [1] beta1  [2] beta2  [3] eps  [4] weight_decay  [5] step (float, advanced here)
//        [6] lr / (1 - beta1^t)   [7] 1 / sqrt(1 - beta2^t)        (outputs of this kernel)
__global__ void adamw_prepare_kernel(float* __restrict__ hyper) {
  pdl_gate();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float t = hyper[5] + 1.0f;
  hyper[5] = t;
  const float bc1 = 1.0f - powf(hyper[1], t);
  const float bc2 = 1.0f - powf(hyper[2], t);
  hyper[6] = hyper[0] / bc1;
  hyper[7] = 1.0f / sqrtf(bc2);
}

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float lr_wd, float b1c, float b2,
                                          float b2c, float step_size, float inv_bc2_sqrt, float eps) {
  p = p - lr_wd * p;
  m = m + b1c * (g - m);
  v = b2 * v + b2c * g * g;
  const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(256)
adamw_multi_kernel(const AdamEntry* __restrict__ table, int n_entries, const float* __restrict__ hyper) {
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
  const AdamEntry e = table[s_e];
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float step_size = hyper[6], inv_bc2_sqrt = hyper[7];
  const float lr_wd = lr * wd, b1c = 1.0f - b1, b2c = 1.0f - b2;
  const size_t base = size_t(blockIdx.x - e.first_block) * kAdamElemsPerBlock;
  if (e.vec) {
    float4 p4[4], g4[4], m4[4], v4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t i = base + size_t(u * 256 + threadIdx.x) * 4;
      if (i < e.n) {
        p4[u] = *reinterpret_cast<const float4*>(e.p + i);
        g4[u] = *reinterpret_cast<const float4*>(e.g + i);
        m4[u] = *reinterpret_cast<const float4*>(e.m + i);
        v4[u] = *reinterpret_cast<const float4*>(e.v + i);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t i = base + size_t(u * 256 + threadIdx.x) * 4;
      if (i < e.n) {
        adamw_one(p4[u].x, g4[u].x, m4[u].x, v4[u].x, lr_wd, b1c, b2, b2c, step_size, inv_bc2_sqrt, eps);
        adamw_one(p4[u].y, g4[u].y, m4[u].y, v4[u].y, lr_wd, b1c, b2, b2c, step_size, inv_bc2_sqrt, eps);
        adamw_one(p4[u].z, g4[u].z, m4[u].z, v4[u].z, lr_wd, b1c, b2, b2c, step_size, inv_bc2_sqrt, eps);
        adamw_one(p4[u].w, g4[u].w, m4[u].w, v4[u].w, lr_wd, b1c, b2, b2c, step_size, inv_bc2_sqrt, eps);
        *reinterpret_cast<float4*>(e.p + i) = p4[u];
        *reinterpret_cast<float4*>(e.m + i) = m4[u];
        *reinterpret_cast<float4*>(e.v + i) = v4[u];
        if (e.w16 != nullptr) {
          uint2 o;
          o.x = pack_bf16(p4[u].x, p4[u].y);
          o.y = pack_bf16(p4[u].z, p4[u].w);
          *reinterpret_cast<uint2*>(e.w16 + i) = o;
        }
      }
    }
  } else {
    for (int u = 0; u < 16; ++u) {
      const size_t i = base + size_t(u * 256 + threadIdx.x);
      if (i < e.n) {
        float p = e.p[i], m = e.m[i], v = e.v[i];
        adamw_one(p, e.g[i], m, v, lr_wd, b1c, b2, b2c, step_size, inv_bc2_sqrt, eps);
        e.p[i] = p; e.m[i] = m; e.v[i] = v;
        if (e.w16 != nullptr) e.w16[i] = __float2bfloat16(p);
      }
    }
  }
}

// Stochastic weight averaging over a list of tensors in one launch (StochasticWeightAveraging.update_parameters /
// avg_fn, helpers/swa_callback.py:246-268):  p_swa <- p_model                              if n_averaged == 0
//                                             p_swa <- p_swa + (p_model - p_swa)/(n_averaged+1) otherwise
// One 32-byte record per tensor: {const float* src = p_model; float* dst = p_swa; uint64 n = elements;
// uint32 first_block = running sum of ceil(n / 4096); uint32 pad}; 16-byte vector path when n % 4 == 0 and both
// pointers are 16-byte aligned, scalar path otherwise.
struct SwaEntry {
  const float* src;
  float* dst;
  unsigned long long n;          // elements
  unsigned int first_block, pad;
};
static_assert(sizeof(SwaEntry) == 32, "SwaEntry layout is part of the C ABI");
constexpr int kSwaElemsPerBlock = 4096;

__global__ void __launch_bounds__(256)
swa_multi_kernel(const SwaEntry* __restrict__ table, int n_entries, float inv_np1, int first) {
  __shared__ int s_e;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_entries - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (table[mid].first_block <= blockIdx.x) lo = mid; else hi = mid - 1;
    }
    s_e = lo;
  }
  __syncthreads();
  const SwaEntry e = table[s_e];
  const size_t base = size_t(blockIdx.x - e.first_block) * kSwaElemsPerBlock;
  const bool vec = (e.n % 4 == 0) && ((reinterpret_cast<uintptr_t>(e.src) | reinterpret_cast<uintptr_t>(e.dst)) % 16 == 0);
  if (vec) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t i = base + size_t(u * 256 + threadIdx.x) * 4;
      if (i < e.n) {
        const float4 m = *reinterpret_cast<const float4*>(e.src + i);
        float4 a = *reinterpret_cast<const float4*>(e.dst + i);
        if (first) a = m;
        else { a.x += (m.x - a.x) * inv_np1; a.y += (m.y - a.y) * inv_np1; a.z += (m.z - a.z) * inv_np1; a.w += (m.w - a.w) * inv_np1; }
        *reinterpret_cast<float4*>(e.dst + i) = a;
      }
    }
  } else {
    for (int u = 0; u < 16; ++u) {
      const size_t i = base + size_t(u * 256 + threadIdx.x);
      if (i < e.n) e.dst[i] = first ? e.src[i] : e.dst[i] + (e.src[i] - e.dst[i]) * inv_np1;
    }
  }
}

}  // namespace pb

extern "C" {

// table: device array of n_entries 32-byte records {const float* p_model; float* p_swa; uint64 n; uint32 first_block;
// uint32 pad}, first_block = running sum of ceil(n / 4096); n_averaged = models averaged so far (0: plain copy).
int passt_swa_update(const void* table, int n_entries, int total_blocks, long long n_averaged, void* stream) {
  using namespace pb;
  if (table == nullptr || n_entries <= 0 || total_blocks <= 0 || n_averaged < 0) return PB_ERR_BAD_ARG;
  swa_multi_kernel<<<total_blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const SwaEntry*>(table), n_entries, 1.0f / float(n_averaged + 1), n_averaged == 0 ? 1 : 0);
  PB_LAUNCH_CHECK();
  return 0;
}

// table: device array of n_entries 64-byte records {float* p; const float* g; float* m; float* v; void* w16_or_null;
// uint64 n; uint32 first_block; uint32 vec; uint64 reserved}; total_blocks = sum of ceil(n / 4096).
// hyper: device float[8] = {lr, beta1, beta2, eps, weight_decay, step, -, -}; step is advanced by one per call.
int passt_adamw_step(const void* table, int n_entries, int total_blocks, float* hyper, void* stream) {
  using namespace pb;
  if (table == nullptr || hyper == nullptr || n_entries <= 0 || total_blocks <= 0) return PB_ERR_BAD_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PB_LAUNCH(adamw_prepare_kernel, 1, 32, 0, st, hyper);
  PB_LAUNCH(adamw_multi_kernel, total_blocks, 256, 0, st, reinterpret_cast<const AdamEntry*>(table), n_entries,
            (const float*)hyper);
  return 0;
}

}  // extern "C"
