// passt_b200 — fused training losses and validation-side post-processing on the logits (sm_100a).
//
// These replace chains of small ATen kernels around the network output; each is ONE launch (latency-bound, a few KB):
//   loss_bce : mean BCE-with-logits against (optionally mixed) multi-hot targets + d loss / d logits
//              (training_step, ex_audioset.py:172-192: y_mix = y*lam + y[perm]*(1-lam); F.binary_cross_entropy_with_logits)
//   loss_ce  : mean cross entropy against (optionally mixed) class indices + d loss / d logits
//              (ex_esc50.py:151-169: CE(y_hat, y)*lam + CE(y_hat, y[perm])*(1-lam))
//   scale_dev: out = in * s[0], s in device memory (backward of the two losses: dlogits * upstream gradient)
//   ens_sigmoid: out = mean_k sigmoid(logits_k) or sigmoid(mean_k logits_k)  (validation_step torch.sigmoid(y_hat),
//              ex_audioset.py:236-238; EnsembelerModel logit averaging, models/passt.py:1021-1036)
// Losses are reduced deterministically: per-clip partials, then the last CTA to finish sums them in index order.
#include "common.cuh"

namespace pb {

__device__ __forceinline__ float block_sum256(float v, float* scratch) {
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
__device__ __forceinline__ float block_max256(float v, float* scratch) {
  v = warp_max(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float s = scratch[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) s = fmaxf(s, scratch[w]);
  return s;
}

// workspace: float partial[B] followed by one unsigned ticket counter (self-resetting)
__device__ __forceinline__ void finish_mean(float clip_sum, float* partial, unsigned* ticket, float* loss, int B,
                                            float inv_count) {
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = clip_sum;
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    s_last = (t == unsigned(B) - 1u);
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence();
    float s = 0.f;
    for (int i = threadIdx.x; i < B; i += 32) s += reinterpret_cast<volatile float*>(partial)[i];
    s = warp_sum(s);                       // fixed order for a given B: run-to-run deterministic
    if (threadIdx.x == 0) {
      loss[0] = s * inv_count;
      *ticket = 0u;
    }
  }
}

// one CTA (256 threads) per clip
__global__ void __launch_bounds__(256)
loss_bce_kernel(const float* __restrict__ logits, const float* __restrict__ target, const int* __restrict__ perm,
                const float* __restrict__ lam, float* __restrict__ loss, float* __restrict__ dlogits,
                float* partial, unsigned* ticket, int B, int C) {
  pdl_gate();
  __shared__ float scratch[8];
  const int b = blockIdx.x;
  const float inv_count = 1.0f / (float(B) * float(C));
  const float* z = logits + size_t(b) * C;
  const float* t0 = target + size_t(b) * C;
  const float* t1 = perm ? target + size_t(perm[b]) * C : nullptr;
  const float l = lam ? lam[b] : 1.0f;
  float acc = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float x = z[c];
    float t = t0[c];
    if (t1) t = t * l + t1[c] * (1.0f - l);
    const float e = expf(-fabsf(x));
    acc += fmaxf(x, 0.f) - x * t + log1pf(e);
    const float sig = x >= 0.f ? 1.0f / (1.0f + e) : e / (1.0f + e);
    if (dlogits) dlogits[size_t(b) * C + c] = (sig - t) * inv_count;
  }
  acc = block_sum256(acc, scratch);
  finish_mean(acc, partial, ticket, loss, B, inv_count);
}

__global__ void __launch_bounds__(256)
loss_ce_kernel(const float* __restrict__ logits, const long long* __restrict__ target, const int* __restrict__ perm,
               const float* __restrict__ lam, float* __restrict__ loss, float* __restrict__ dlogits, float* partial,
               unsigned* ticket, int B, int C) {
  pdl_gate();
  __shared__ float scratch[8];
  const int b = blockIdx.x;
  const float* z = logits + size_t(b) * C;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += 256) mx = fmaxf(mx, z[c]);
  mx = block_max256(mx, scratch);
  float se = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) se += expf(z[c] - mx);
  se = block_sum256(se, scratch);
  const float lse = mx + logf(se);
  const int y0 = int(target[b]);
  const int y1 = perm ? int(target[perm[b]]) : y0;
  const float l = lam ? lam[b] : 1.0f;
  const float inv_b = 1.0f / float(B);
  if (dlogits) {
    for (int c = threadIdx.x; c < C; c += 256) {
      float g = expf(z[c] - lse);
      if (c == y0) g -= l;
      if (c == y1) g -= (1.0f - l);
      dlogits[size_t(b) * C + c] = g * inv_b;
    }
  }
  // same association as the reference: CE(z, y)*lam + CE(z, y[perm])*(1-lam)
  const float clip = (lse - z[y0]) * l + (lse - z[y1]) * (1.0f - l);
  finish_mean(clip, partial, ticket, loss, B, inv_b);
}

__global__ void __launch_bounds__(256)
scale_dev_kernel(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ s, size_t n) {
  pdl_gate();
  const size_t i = size_t(blockIdx.x) * 256 + threadIdx.x;
  if (i < n) out[i] = in[i] * s[0];
}

// ptrs[k]: logits of net k, [n] floats each.  mode 0: sigmoid(mean_k logits_k) (EnsembelerModel then sigmoid);
// mode 1: mean_k sigmoid(logits_k).  K == 1: plain sigmoid.
struct EnsPtrs { const float* p[16]; };
__global__ void __launch_bounds__(256)
ens_sigmoid_kernel(EnsPtrs ptrs, int K, float* __restrict__ mean_logits, float* __restrict__ prob, size_t n, int mode) {
  const size_t i = size_t(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  float sl = 0.f, sp = 0.f;
  for (int k = 0; k < K; ++k) {
    const float x = ptrs.p[k][i];
    sl += x;
    sp += 1.0f / (1.0f + expf(-x));
  }
  const float ml = sl / float(K);
  if (mean_logits) mean_logits[i] = ml;
  if (prob) prob[i] = mode == 0 ? 1.0f / (1.0f + expf(-ml)) : sp / float(K);
}

// Average precision per class, sklearn.metrics.average_precision_score(average=None) semantics including tied scores
// (validation_epoch_end, ex_audioset.py:262-266):  AP_c = 1/P_c * sum over positives i of  #pos(s >= s_i) / #all(s >= s_i)
// -- every positive sharing a threshold value contributes that threshold's precision once, which is exactly
// sum_n (R_n - R_{n-1}) P_n over the distinct thresholds.  O(P_c * n) comparisons per class, no sort, no host hop.
// One CTA per class; scores / targets are [n, C] row-major (the layout validation_step concatenates).  A class
// without positives yields NaN.
__global__ void __launch_bounds__(256)
average_precision_kernel(const float* __restrict__ scores, const float* __restrict__ targets, float* __restrict__ ap,
                         int n, int C) {
  extern __shared__ float s_sc[];            // this class's scores; sign bit of the copy is NOT used: labels go to s_lb
  unsigned char* s_lb = reinterpret_cast<unsigned char*>(s_sc + n);
  __shared__ float s_red[8];
  __shared__ int s_cnt[8];
  const int c = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int npos = 0;
  for (int i = threadIdx.x; i < n; i += 256) {
    s_sc[i] = scores[size_t(i) * C + c];
    const unsigned char l = targets[size_t(i) * C + c] > 0.5f ? 1 : 0;
    s_lb[i] = l;
    npos += l;
  }
  npos = __reduce_add_sync(0xffffffffu, npos);
  if (lane == 0) s_cnt[warp] = npos;
  __syncthreads();
  int P = 0;
  for (int w = 0; w < 8; ++w) P += s_cnt[w];
  float acc = 0.f;
  // warp w takes the positives whose sample index i satisfies (i / 32) % 8 == w; lanes sweep the n samples
  for (int i0 = warp * 32; i0 < n; i0 += 256) {
    for (int k = 0; k < 32 && i0 + k < n; ++k) {
      const int i = i0 + k;
      if (!s_lb[i]) continue;                // warp-uniform (same i for all lanes)
      const float si = s_sc[i];
      int ge = 0, gep = 0;
      for (int j = lane; j < n; j += 32) {
        const bool hit = s_sc[j] >= si;
        ge += hit;
        gep += hit && s_lb[j];
      }
      ge = __reduce_add_sync(0xffffffffu, ge);
      gep = __reduce_add_sync(0xffffffffu, gep);
      acc += float(gep) / float(ge);
    }
  }
  if (lane == 0) s_red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    ap[c] = P > 0 ? t / float(P) : __int_as_float(0x7fc00000);
  }
}

}  // namespace pb

extern "C" {

// scores, targets: device f32 [n, C]; ap: device f32 [C].  n is limited by shared memory (5 bytes per sample).
int passt_average_precision(const float* scores, const float* targets, float* ap, int n, int C, void* stream) {
  using namespace pb;
  if (!scores || !targets || !ap || n <= 0 || C <= 0) return PB_ERR_BAD_ARG;
  const size_t smem = size_t(n) * 5 + 16;
  if (smem > 220 * 1024) return PB_ERR_BAD_ARG;
  PB_SET_SMEM_ONCE(220 * 1024, average_precision_kernel);
  average_precision_kernel<<<C, 256, smem, (cudaStream_t)stream>>>(scores, targets, ap, n, C);
  PB_LAUNCH_CHECK();
  return 0;
}

size_t passt_loss_workspace_bytes(int B) { return size_t(B < 1 ? 1 : B) * 4 + 64; }

static inline unsigned* ticket_of(void* ws, int B) {
  return reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(ws) + ((size_t(B) * 4 + 15) / 16) * 16);
}

// workspace must be zero-initialised once by the caller (the ticket counter resets itself after every launch)
int passt_loss_bce(const float* logits, const float* target, const int* perm, const float* lam, float* loss,
                   float* dlogits, void* workspace, int B, int C, void* stream) {
  using namespace pb;
  if (!logits || !target || !loss || !workspace || B <= 0 || C <= 0) return PB_ERR_BAD_ARG;
  if ((perm == nullptr) != (lam == nullptr)) return PB_ERR_BAD_ARG;
  PB_LAUNCH(loss_bce_kernel, B, 256, 0, (cudaStream_t)stream, logits, target, perm, lam, loss, dlogits,
            reinterpret_cast<float*>(workspace), ticket_of(workspace, B), B, C);
  return 0;
}

int passt_loss_ce(const float* logits, const long long* target, const int* perm, const float* lam, float* loss,
                  float* dlogits, void* workspace, int B, int C, void* stream) {
  using namespace pb;
  if (!logits || !target || !loss || !workspace || B <= 0 || C <= 0) return PB_ERR_BAD_ARG;
  if ((perm == nullptr) != (lam == nullptr)) return PB_ERR_BAD_ARG;
  PB_LAUNCH(loss_ce_kernel, B, 256, 0, (cudaStream_t)stream, logits, target, perm, lam, loss, dlogits,
            reinterpret_cast<float*>(workspace), ticket_of(workspace, B), B, C);
  return 0;
}

int passt_scale_dev(float* out, const float* in, const float* scalar_dev, size_t n, void* stream) {
  using namespace pb;
  if (!out || !in || !scalar_dev) return PB_ERR_BAD_ARG;
  if (n == 0) return 0;
  PB_LAUNCH(scale_dev_kernel, unsigned((n + 255) / 256), 256, 0, (cudaStream_t)stream, out, in, scalar_dev, n);
  return 0;
}

// logits_ptrs: HOST array of K device pointers (K <= 16), each [n] floats
int passt_ens_sigmoid(const float* const* logits_ptrs, int K, float* mean_logits, float* prob, size_t n, int mode,
                      void* stream) {
  using namespace pb;
  if (!logits_ptrs || K <= 0 || K > 16 || (mode != 0 && mode != 1)) return PB_ERR_BAD_ARG;
  if (n == 0) return 0;
  EnsPtrs pp;
  for (int k = 0; k < 16; ++k) pp.p[k] = k < K ? logits_ptrs[k] : nullptr;
  ens_sigmoid_kernel<<<unsigned((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pp, K, mean_logits, prob, n, mode);
  PB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
