// passt_b200 — waveform-side augmentation on the device, in the pass that stages a batch for the mel kernel (sm_100a).
//
// Replaces the per-sample CPU work of the reference's loader workers:
//   pydub_augment gain      waveform * 10^(gain_dB/20)                         audioset/dataset.py:107-115
//   pad_or_truncate         zero-pad / cut to clip_length                      audioset/dataset.py:315-320
//   roll                    x.roll(shift, axis=1)                              audioset/dataset.py:323-339 (shift in [-50, 50])
//   MixupDataset            x1 -= mean; x2 -= mean; x = x1*l + x2*(1-l); x -= mean; y = y1*l + y2*(1-l)
//                                                                              audioset/dataset.py:118-140
// in that order per clip.  The partner of a mixed clip is another clip of the SAME batch (after its own gain / pad /
// roll), the device-friendly equivalent of the reference's "another random dataset item".
//
// One thread-block CLUSTER of 8 CTAs owns one clip (8 x B CTAs fill the 148 SMs at the bench batch); the three global
// reductions a mixed clip needs (mean of each source, mean of the mixture) go through distributed shared memory.
// HBM traffic: read each source once (twice for mixed clips: the partner), write the staged batch once; the mixture's
// mean is subtracted in a second pass over the CTA's own L/8 slice, which is still in L2.
#include "common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace pb {

constexpr int kAugCluster = 8;
constexpr int kAugThreads = 512;

struct WaveAugParams {
  const float* raw;            // all source clips
  const long long* src_off;    // [B] element offset of clip b in raw
  const int* src_len;          // [B] samples available (longer clips are truncated to L, shorter ones zero-padded)
  const float* gain;           // [B] linear amplitude or nullptr (1.0)
  const int* shift;            // [B] roll shift or nullptr (0)
  const int* mix_idx;          // [B] partner clip or -1, or nullptr (no waveform mixup)
  const float* mix_lam;        // [B]
  float* out;                  // [B, L]
  const float* tgt;            // [B, C] targets or nullptr
  float* tgt_out;              // [B, C]
  int B, L, C;
};

__device__ __forceinline__ float aug_sample(const WaveAugParams& p, int b, int i, float g, int s) {
  int j = i - s;
  j %= p.L;
  if (j < 0) j += p.L;
  return j < p.src_len[b] ? g * __ldg(p.raw + p.src_off[b] + j) : 0.f;
}

// sum over the cluster of one float per thread; every thread of every CTA gets the total
__device__ __forceinline__ float cluster_sum(cg::cluster_group& cluster, float v, float* s_part, float* s_cta) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kAugThreads / 32; ++w) t += s_part[w];
    *s_cta = t;
  }
  cluster.sync();                       // every CTA's partial is visible cluster-wide
  float total = 0.f;
  for (int r = 0; r < kAugCluster; ++r) total += *cluster.map_shared_rank(s_cta, r);   // fixed order: deterministic
  cluster.sync();                       // nobody overwrites s_cta before all ranks have read it
  return total;
}

__global__ void __cluster_dims__(kAugCluster, 1, 1) __launch_bounds__(kAugThreads)
wave_augment_kernel(const WaveAugParams p) {
  __shared__ float s_part[kAugThreads / 32];
  __shared__ float s_cta;
  cg::cluster_group cluster = cg::this_cluster();
  const int b = blockIdx.x / kAugCluster;
  const int rank = int(cluster.block_rank());
  const int per = (p.L + kAugCluster - 1) / kAugCluster;
  const int i0 = rank * per, i1 = min(p.L, i0 + per);
  const float g = p.gain ? p.gain[b] : 1.0f;
  const int s = p.shift ? p.shift[b] : 0;
  const int partner = p.mix_idx ? p.mix_idx[b] : -1;
  float* out = p.out + size_t(b) * p.L;

  if (partner < 0) {
    for (int i = i0 + threadIdx.x; i < i1; i += kAugThreads) out[i] = aug_sample(p, b, i, g, s);
    if (p.tgt && rank == 0)
      for (int c = threadIdx.x; c < p.C; c += kAugThreads) p.tgt_out[size_t(b) * p.C + c] = p.tgt[size_t(b) * p.C + c];
    return;                               // block-uniform across the cluster (same b -> same partner)
  }
  const float g2 = p.gain ? p.gain[partner] : 1.0f;
  const int s2 = p.shift ? p.shift[partner] : 0;
  const float l = p.mix_lam[b];
  // ---- means of the two (gained, padded, rolled) sources
  float a1 = 0.f, a2 = 0.f;
  for (int i = i0 + threadIdx.x; i < i1; i += kAugThreads) {
    a1 += aug_sample(p, b, i, g, s);
    a2 += aug_sample(p, partner, i, g2, s2);
  }
  const float m1 = cluster_sum(cluster, a1, s_part, &s_cta) / float(p.L);
  const float m2 = cluster_sum(cluster, a2, s_part, &s_cta) / float(p.L);
  // ---- mixture, written once; its mean is removed in a second pass over this CTA's slice (L2-resident)
  float ax = 0.f;
  for (int i = i0 + threadIdx.x; i < i1; i += kAugThreads) {
    const float x = (aug_sample(p, b, i, g, s) - m1) * l + (aug_sample(p, partner, i, g2, s2) - m2) * (1.0f - l);
    out[i] = x;
    ax += x;
  }
  const float mx = cluster_sum(cluster, ax, s_part, &s_cta) / float(p.L);
  for (int i = i0 + threadIdx.x; i < i1; i += kAugThreads) out[i] -= mx;   // same thread wrote out[i]: no fence needed
  if (p.tgt && rank == 0)
    for (int c = threadIdx.x; c < p.C; c += kAugThreads)
      p.tgt_out[size_t(b) * p.C + c] = p.tgt[size_t(b) * p.C + c] * l + p.tgt[size_t(partner) * p.C + c] * (1.0f - l);
}

}  // namespace pb

extern "C" {

// raw: device f32, all source clips; src_off [B] int64 element offsets; src_len [B] int32; gain [B] f32 or NULL;
// shift [B] int32 or NULL; mix_idx [B] int32 (-1: not mixed) + mix_lam [B] f32, or both NULL; out [B, L] f32;
// tgt / tgt_out [B, C] f32 or both NULL (targets of mixed clips are mixed with the same lam).
int passt_wave_augment(const float* raw, const long long* src_off, const int* src_len, const float* gain,
                       const int* shift, const int* mix_idx, const float* mix_lam, float* out, const float* tgt,
                       float* tgt_out, int B, int L, int C, void* stream) {
  using namespace pb;
  if (!raw || !src_off || !src_len || !out || B <= 0 || L <= 0) return PB_ERR_BAD_ARG;
  if ((mix_idx == nullptr) != (mix_lam == nullptr) || (tgt == nullptr) != (tgt_out == nullptr)) return PB_ERR_BAD_ARG;
  if (tgt && C <= 0) return PB_ERR_BAD_ARG;
  WaveAugParams p{raw, src_off, src_len, gain, shift, mix_idx, mix_lam, out, tgt, tgt_out, B, L, C};
  wave_augment_kernel<<<B * kAugCluster, kAugThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  PB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
