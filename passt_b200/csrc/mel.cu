// passt_b200 — fused waveform -> normalised log-mel frontend (sm_100a).
//
// Replaces AugmentMelSTFT.forward (reference models/preprocess.py:57-86):
//   pre-emphasis conv1d [-0.97, 1] (:59) -> torch.stft(n_fft=1024, hop=320, win=hann(800), center/reflect) (:60-61)
//   -> power (:62) -> kaldi mel banks for (fmin,fmax) (:71-74, torchaudio/compliance/kaldi.py:436-511)
//   -> matmul (:76) -> log(x+1e-5) (:78) -> FrequencyMasking/TimeMasking iid (:80-82,
//   torchaudio/functional/functional.py:813-882) -> (x+4.5)/5 (:84)
// in ONE kernel: the windowed frames, the complex spectrum and the dense filterbank never exist in HBM.
//
// Work decomposition: one CTA = one clip x 32 consecutive frames; one warp = one frame at a time.
// The 1024-point real FFT is a 512-point complex FFT (even/odd packing) done as three radix-8 Stockham
// passes in a per-warp shared-memory buffer, followed by the real-FFT untangling butterfly.
#include "common.cuh"
#include <math.h>

namespace pb {

constexpr int kNfft = 1024;
constexpr int kHalf = 512;
constexpr int kMelBins = 128;
constexpr int kMaxW = 32;        // max non-zeros per triangular filter (widest PaSST filter: 28 taps at fmax = 16 kHz)
constexpr int kMelWarps = 12;
constexpr int kFramesPerRound = 48;   // 4 frames per warp, staged and stored as 48-frame rows
constexpr int kRoundsPerCta = 2;
constexpr int kFramesPerCta = kFramesPerRound * kRoundsPerCta;
constexpr int kFftPad = kHalf + kHalf / 8;   // padded per-warp FFT buffer: phys(i) = i + (i >> 3)

struct MelTables {
  float2 tw1024[1024];  // exp(-2*pi*i*m/1024)
  float win[1024];      // hann(win_length, periodic=False) centred in n_fft, zero elsewhere
  int win_lo, win_hi;   // non-zero support [win_lo, win_hi)
};

// sparse filterbank, rebuilt on device whenever (fmin, fmax) change
struct MelBank {
  float wT[kMaxW * kMelBins];  // wT[j*128 + m] = weight of fft bin lo[m]+j for mel bin m
  int lo[kMelBins];
  int cnt[kMelBins];
  int overflow;  // set if some filter has more than kMaxW taps
};

__global__ void mel_tables_kernel(MelTables* t, int win_length) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 1024) {
    double s, c;
    sincospi(-2.0 * double(i) / 1024.0, &s, &c);
    t->tw1024[i] = make_float2(float(c), float(s));
    const int lo = (kNfft - win_length) / 2;
    float w = 0.f;
    if (i >= lo && i < lo + win_length) {
      // torch.hann_window(N, periodic=False): 0.5 - 0.5*cos(2*pi*n/(N-1))
      const double n = double(i - lo);
      w = float(0.5 - 0.5 * cospi(2.0 * n / double(win_length - 1)));
    }
    t->win[i] = w;
    if (i == 0) { t->win_lo = lo; t->win_hi = lo + win_length; }
  }
}

// kaldi get_mel_banks (torchaudio/compliance/kaldi.py:436-511), vtln_warp_factor == 1.0 branch, fp32 like torch.
// mel endpoints come in as doubles computed the way the python scalars are (mel_scale_scalar uses math.log).
__global__ void mel_bank_kernel(MelBank* bank, double mel_low, double mel_high, float sample_rate,
                                const double* __restrict__ band_dev) {
  pdl_gate();
  __shared__ float melk[kHalf];
  const int tid = threadIdx.x;
  if (band_dev != nullptr) {   // (fmin, fmax) live in device memory (CUDA-graph replays with per-step augmentation)
    mel_low = 1127.0 * log(1.0 + band_dev[0] / 700.0);
    mel_high = 1127.0 * log(1.0 + band_dev[1] / 700.0);
  }
  const float bin_width = sample_rate / float(kNfft);
  for (int k = tid; k < kHalf; k += blockDim.x) {
    const float f = __fmul_rn(bin_width, float(k));
    melk[k] = __fmul_rn(1127.0f, logf(__fadd_rn(1.0f, __fdiv_rn(f, 700.0f))));
  }
  if (tid == 0) bank->overflow = 0;
  __syncthreads();
  if (tid < kMelBins) {
    const float delta = float((mel_high - mel_low) / double(kMelBins + 1));
    const float lowf = float(mel_low);
    const float left = __fadd_rn(lowf, __fmul_rn(float(tid), delta));
    const float center = __fadd_rn(lowf, __fmul_rn(float(tid) + 1.0f, delta));
    const float right = __fadd_rn(lowf, __fmul_rn(float(tid) + 2.0f, delta));
    const float inv_up = __fsub_rn(center, left);
    const float inv_dn = __fsub_rn(right, center);
    int lo = -1, cnt = 0;
    for (int k = 0; k < kHalf; ++k) {
      const float up = __fdiv_rn(__fsub_rn(melk[k], left), inv_up);
      const float dn = __fdiv_rn(__fsub_rn(right, melk[k]), inv_dn);
      const float w = fmaxf(0.f, fminf(up, dn));
      if (w > 0.f) {
        if (lo < 0) lo = k;
        const int j = k - lo;
        if (j < kMaxW) bank->wT[j * kMelBins + tid] = w;
        cnt = j + 1;
      }
    }
    if (lo < 0) { lo = 0; cnt = 0; }
    if (cnt > kMaxW) { bank->overflow = 1; cnt = kMaxW; }
    for (int j = cnt; j < kMaxW; ++j) bank->wT[j * kMelBins + tid] = 0.f;
    // interior zeros between lo and lo+cnt keep weight 0 — fill them explicitly
    for (int j = 0; j < cnt; ++j) {
      const int k = lo + j;
      const float up = __fdiv_rn(__fsub_rn(melk[k], left), inv_up);
      const float dn = __fdiv_rn(__fsub_rn(right, melk[k]), inv_dn);
      bank->wT[j * kMelBins + tid] = fmaxf(0.f, fminf(up, dn));
    }
    bank->lo[tid] = lo;
    bank->cnt[tid] = cnt;
  }
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward-DFT quarter turn)
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

// in-register 8-point forward DFT, natural order in / natural order out
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
  const float r = 0.70710678118654752f;
  // stage 1: pairs (k, k+4)
  float2 a0 = cadd(v[0], v[4]), a4 = csub(v[0], v[4]);
  float2 a1 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
  float2 a2 = cadd(v[2], v[6]), a6 = csub(v[2], v[6]);
  float2 a3 = cadd(v[3], v[7]), a7 = csub(v[3], v[7]);
  // twiddles W8^k on the odd half: W8^1 = r(1 - i), W8^2 = -i, W8^3 = r(-1 - i)
  a5 = make_float2(r * (a5.x + a5.y), r * (a5.y - a5.x));
  a6 = mul_mi(a6);
  a7 = make_float2(r * (a7.y - a7.x), r * (-a7.x - a7.y));
  // stage 2 on evens (a0..a3) and odds (a4..a7): 4-point DFTs
  float2 b0 = cadd(a0, a2), b2 = csub(a0, a2);
  float2 b1 = cadd(a1, a3), b3 = mul_mi(csub(a1, a3));
  float2 b4 = cadd(a4, a6), b6 = csub(a4, a6);
  float2 b5 = cadd(a5, a7), b7 = mul_mi(csub(a5, a7));
  // stage 3
  v[0] = cadd(b0, b1); v[4] = csub(b0, b1);
  v[2] = cadd(b2, b3); v[6] = csub(b2, b3);
  v[1] = cadd(b4, b5); v[5] = csub(b4, b5);
  v[3] = cadd(b6, b7); v[7] = csub(b6, b7);
}

__device__ __forceinline__ int fphys(int i) { return i + (i >> 3); }   // bank-conflict-free padding of the FFT buffer

// one Stockham radix-8 pass over a 512-point buffer held by one warp (in place via registers).
// tw: per-pass twiddle table laid out [r-1][k] (k contiguous) so that lanes read consecutive addresses.
template <int Ns>
__device__ __forceinline__ void radix8_pass(float2* buf, const float2* tw, int lane) {
  float2 v[2][8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
#pragma unroll
    for (int r = 0; r < 8; ++r) v[h][r] = buf[fphys(j + 64 * r)];
    if (Ns > 1) {
      const int k = j % Ns;
#pragma unroll
      for (int r = 1; r < 8; ++r) v[h][r] = cmul(v[h][r], tw[(r - 1) * Ns + k]);   // exp(-2*pi*i*r*k/(8*Ns))
    }
    dft8(v[h]);
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
    const int base = (j / Ns) * Ns * 8 + (j % Ns);
#pragma unroll
    for (int r = 0; r < 8; ++r) buf[fphys(base + r * Ns)] = v[h][r];
  }
  __syncwarp();
}

struct MelParams {
  const float* wave;  // [B, L]
  float* out;         // [B, 128, T]
  int B, L, T, hop;
  const float* rnd;   // [4, B] uniform draws (freq value, freq min, time value, time min) or nullptr (eval)
  int freqm, timem;   // SpecAugment mask params (0 = off)
  float preemph;      // 0.97
};

struct MelSmem {
  static constexpr int kTw = 0;                                  // 512 float2: exp(-2 pi i k / 1024), untangle
  static constexpr int kT2 = kTw + 512 * 8;                      // 7 x 8  float2: pass-2 twiddles
  static constexpr int kT3 = kT2 + 7 * 8 * 8;                    // 7 x 64 float2: pass-3 twiddles
  static constexpr int kWin = kT3 + 7 * 64 * 8;                  // 1024 floats
  static constexpr int kWT = kWin + 1024 * 4;                    // kMaxW x 128 floats
  static constexpr int kLo = kWT + kMaxW * kMelBins * 4;         // 128 ints
  static constexpr int kCnt = kLo + kMelBins * 4;
  static constexpr int kOut = kCnt + kMelBins * 4;               // 128 x (kFramesPerRound + 1) floats
  static constexpr int kFft = kOut + kMelBins * (kFramesPerRound + 1) * 4;
  static constexpr int kTotal = kFft + kMelWarps * kFftPad * 8;
};

__global__ void __launch_bounds__(kMelWarps * 32, 2)
mel_kernel(const MelParams p, const MelTables* __restrict__ tabs, const MelBank* __restrict__ bank) {
  pdl_gate();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float2* s_tw = reinterpret_cast<float2*>(smem_raw + MelSmem::kTw);
  float2* s_t2 = reinterpret_cast<float2*>(smem_raw + MelSmem::kT2);
  float2* s_t3 = reinterpret_cast<float2*>(smem_raw + MelSmem::kT3);
  float* s_win = reinterpret_cast<float*>(smem_raw + MelSmem::kWin);
  float* s_wT = reinterpret_cast<float*>(smem_raw + MelSmem::kWT);
  int* s_lo = reinterpret_cast<int*>(smem_raw + MelSmem::kLo);
  int* s_cnt = reinterpret_cast<int*>(smem_raw + MelSmem::kCnt);
  float* s_out = reinterpret_cast<float*>(smem_raw + MelSmem::kOut);
  float2* s_fft = reinterpret_cast<float2*>(smem_raw + MelSmem::kFft);
  constexpr int kOutLd = kFramesPerRound + 1;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;

  for (int i = tid; i < 512; i += blockDim.x) s_tw[i] = tabs->tw1024[i];
  for (int i = tid; i < 7 * 8; i += blockDim.x) s_t2[i] = tabs->tw1024[(i / 8 + 1) * (i % 8) * 16];
  for (int i = tid; i < 7 * 64; i += blockDim.x) s_t3[i] = tabs->tw1024[(i / 64 + 1) * (i % 64) * 2];
  for (int i = tid; i < 1024; i += blockDim.x) s_win[i] = tabs->win[i];
  for (int i = tid; i < kMaxW * kMelBins; i += blockDim.x) s_wT[i] = bank->wT[i];
  if (tid < kMelBins) { s_lo[tid] = bank->lo[tid]; s_cnt[tid] = bank->cnt[tid]; }
  const bool bank_overflow = bank->overflow != 0;   // a filter wider than kMaxW taps: poison the output, never truncate silently
  __syncthreads();

  const float* x = p.wave + size_t(b) * p.L;
  const int Ly = p.L - 1;  // pre-emphasised length
  float2* buf = s_fft + warp * kFftPad;
  float* pw = reinterpret_cast<float*>(buf);  // power spectrum reuses the FFT buffer (512 floats, unpadded)
  const bool vec_ok = ((p.L & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & 7) == 0);

  // per-example SpecAugment bands (torchaudio mask_along_axis_iid: start=floor(min), end=start+floor(value))
  int f_lo = 0, f_hi = 0, m_lo = 0, m_hi = 0;
  if (p.rnd != nullptr) {
    if (p.freqm > 0) {
      const float value = p.rnd[0 * p.B + b] * float(p.freqm);
      const float mn = p.rnd[1 * p.B + b] * (float(kMelBins) - value);
      f_lo = int(mn); f_hi = int(mn) + int(value);
    }
    if (p.timem > 0) {
      const float value = p.rnd[2 * p.B + b] * float(p.timem);
      const float mn = p.rnd[3 * p.B + b] * (float(p.T) - value);
      m_lo = int(mn); m_hi = int(mn) + int(value);
    }
  }

  for (int round = 0; round < kRoundsPerCta; ++round) {
    const int t0 = blockIdx.x * kFramesPerCta + round * kFramesPerRound;
    if (t0 >= p.T) break;   // block-uniform
    for (int fi = warp; fi < kFramesPerRound; fi += kMelWarps) {
      const int t = t0 + fi;
      if (t >= p.T) break;  // warp-uniform
      // ---- load: z[n] = f[2n] + i f[2n+1], f[m] = win[m] * y[t*hop - 512 + m], y = pre-emphasised, reflect pad
      const int j0 = t * p.hop - kNfft / 2;
      const bool interior = vec_ok && j0 >= 0 && (j0 + kNfft + 2) <= Ly && ((j0 & 1) == 0);
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int n = lane + 32 * i;
        const float w0 = s_win[2 * n], w1 = s_win[2 * n + 1];
        float2 z = make_float2(0.f, 0.f);
        if (w0 != 0.f || w1 != 0.f) {
          if (interior) {
            const float2 a = __ldg(reinterpret_cast<const float2*>(x + j0 + 2 * n));
            const float c = __ldg(x + j0 + 2 * n + 2);
            z.x = w0 * (a.y - p.preemph * a.x);
            z.y = w1 * (c - p.preemph * a.y);
          } else {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float w = e ? w1 : w0;
              float val = 0.f;
              if (w != 0.f) {
                int j = j0 + 2 * n + e;
                if (j < 0) j = -j;
                if (j >= Ly) j = 2 * (Ly - 1) - j;
                const float x0 = __ldg(x + j), x1 = __ldg(x + j + 1);
                val = w * (x1 - p.preemph * x0);
              }
              v[e] = val;
            }
            z = make_float2(v[0], v[1]);
          }
        }
        buf[fphys(n)] = z;
      }
      __syncwarp();
      // ---- 512-point complex FFT
      radix8_pass<1>(buf, s_t2, lane);
      radix8_pass<8>(buf, s_t2, lane);
      radix8_pass<64>(buf, s_t3, lane);
      // ---- untangle to the 1024-point real spectrum, power for bins 0..511
      float pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = lane + 32 * i;
        const float2 zk = buf[fphys(k)];
        float2 zc = buf[fphys((kHalf - k) & (kHalf - 1))];
        zc.y = -zc.y;
        const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
        const float2 d = make_float2(0.5f * (zk.x - zc.x), 0.5f * (zk.y - zc.y));
        const float2 o = make_float2(d.y, -d.x);  // -i * d
        const float2 xo = cmul(s_tw[k], o);
        const float re = e.x + xo.x, im = e.y + xo.y;
        pk[i] = re * re + im * im;
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 16; ++i) pw[lane + 32 * i] = pk[i];
      __syncwarp();
      // ---- sparse triangular filterbank + log + masks + affine
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = lane + 32 * i;
        const int lo = s_lo[m], cnt = s_cnt[m];
        float acc = 0.f;
        for (int j = 0; j < cnt; ++j) acc = fmaf(s_wT[j * kMelBins + m], pw[lo + j], acc);
        float v = logf(acc + 1e-5f);
        if ((m >= f_lo && m < f_hi) || (t >= m_lo && t < m_hi)) v = 0.f;
        s_out[m * kOutLd + fi] = bank_overflow ? __int_as_float(0x7fc00000) : (v + 4.5f) / 5.0f;
      }
      __syncwarp();
    }
    __syncthreads();
    // ---- coalesced store: up to 48 consecutive frames per mel row
    const int nt = min(kFramesPerRound, p.T - t0);
    float* o = p.out + size_t(b) * kMelBins * p.T + t0;
    for (int idx = tid; idx < kMelBins * kFramesPerRound; idx += blockDim.x) {
      const int m = idx / kFramesPerRound, f = idx - m * kFramesPerRound;
      if (f < nt) o[size_t(m) * p.T + f] = s_out[m * kOutLd + f];
    }
    __syncthreads();
  }
}

constexpr int kMelSmemBytes = MelSmem::kTotal;

}  // namespace pb

extern "C" {

// Workspace: tables (hann window + twiddles) and the sparse filterbank live in caller-provided device memory.
size_t passt_mel_workspace_bytes() { return sizeof(pb::MelTables) + sizeof(pb::MelBank) + 256; }

// One-time table init for a given win_length (reference: models/preprocess.py:38-40 hann window buffer).
int passt_mel_init(void* workspace, int win_length, void* stream) {
  using namespace pb;
  if (!workspace || win_length <= 0 || win_length > kNfft) return PB_ERR_BAD_ARG;
  MelTables* t = reinterpret_cast<MelTables*>(workspace);
  mel_tables_kernel<<<4, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(t, win_length);
  PB_LAUNCH_CHECK();
  return 0;
}

// Rebuild the sparse kaldi filterbank for (fmin, fmax) on device (no host filterbank, no H2D copy).
int passt_mel_set_band(void* workspace, double fmin, double fmax, int sample_rate, void* stream) {
  using namespace pb;
  if (!workspace) return PB_ERR_BAD_ARG;
  const double nyq = 0.5 * sample_rate;
  if (fmax <= 0.0) fmax += nyq;
  if (!(0.0 <= fmin && fmin < nyq && 0.0 < fmax && fmax <= nyq && fmin < fmax)) return PB_ERR_BAD_ARG;
  MelBank* bank = reinterpret_cast<MelBank*>(reinterpret_cast<uint8_t*>(workspace) +
                                             ((sizeof(MelTables) + 255) / 256) * 256);
  const double mel_low = 1127.0 * log(1.0 + fmin / 700.0);
  const double mel_high = 1127.0 * log(1.0 + fmax / 700.0);
  PB_LAUNCH(mel_bank_kernel, 1, 512, 0, reinterpret_cast<cudaStream_t>(stream), bank, mel_low, mel_high,
            float(sample_rate), (const double*)nullptr);
  return 0;
}

// Same, with (fmin, fmax) read from device memory (double[2]) at kernel run time: the launch is identical every step,
// so it can live inside a captured CUDA graph while the band augmentation still changes per step.
int passt_mel_set_band_dev(void* workspace, const double* band_dev, int sample_rate, void* stream) {
  using namespace pb;
  if (!workspace || !band_dev) return PB_ERR_BAD_ARG;
  MelBank* bank = reinterpret_cast<MelBank*>(reinterpret_cast<uint8_t*>(workspace) +
                                             ((sizeof(MelTables) + 255) / 256) * 256);
  PB_LAUNCH(mel_bank_kernel, 1, 512, 0, reinterpret_cast<cudaStream_t>(stream), bank, 0.0, 0.0, float(sample_rate),
            band_dev);
  return 0;
}

// wave [B, L] fp32 -> out [B, 128, T] fp32, T = 1 + (L-1)/hop. rnd = [4,B] uniforms for SpecAugment or NULL.
int passt_mel_forward(const void* workspace, const float* wave, float* out, int B, int L, int hop,
                      const float* rnd, int freqm, int timem, void* stream) {
  using namespace pb;
  if (!workspace || !wave || !out || B <= 0 || L < kNfft || hop <= 0) return PB_ERR_BAD_ARG;
  const MelTables* t = reinterpret_cast<const MelTables*>(workspace);
  const MelBank* bank = reinterpret_cast<const MelBank*>(reinterpret_cast<const uint8_t*>(workspace) +
                                                         ((sizeof(MelTables) + 255) / 256) * 256);
  MelParams p;
  p.wave = wave; p.out = out; p.B = B; p.L = L; p.hop = hop;
  p.T = 1 + (L - 1) / hop;
  p.rnd = rnd; p.freqm = freqm; p.timem = timem; p.preemph = 0.97f;
  PB_SET_SMEM_ONCE(kMelSmemBytes, mel_kernel);
  dim3 grid((p.T + kFramesPerCta - 1) / kFramesPerCta, B);
  PB_LAUNCH(mel_kernel, grid, kMelWarps * 32, kMelSmemBytes, reinterpret_cast<cudaStream_t>(stream), p, t, bank);
  return 0;
}

}  // extern "C"
