/* passt_b200 — C ABI of the sm_100a kernel library (passt_b200/lib/libpasst_b200.so).
 *
 * The reference (kkoutini/PaSST) has no FFI: its hot path is Python calling torch / torchaudio ops.  Each entry
 * point below replaces the torch op sequence cited next to it; the Python modules in passt_b200/ (drop-ins for
 * models/preprocess.py and models/passt.py) are the only callers.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions: plain device pointers + sizes, no torch types; all calls are asynchronous on `stream`
 * (a cudaStream_t passed as void*); return 0 on success, a positive cudaError_t, or a negative library code
 * (-2 bad argument, -3 driver/tensor-map failure).  bf16 tensors are passed as `void*`.
 */
#ifndef PASST_B200_H
#define PASST_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- frontend: AugmentMelSTFT.forward, reference models/preprocess.py:57-86 --------------------------------- */
/* bytes of device scratch holding the hann/twiddle tables and the sparse mel filterbank */
size_t passt_mel_workspace_bytes(void);
/* hann(win_length, periodic=False) centred in n_fft=1024 + FFT twiddles (preprocess.py:38-40) */
int passt_mel_init(void* workspace, int win_length, void* stream);
/* kaldi triangular filterbank for (fmin, fmax) built ON DEVICE (preprocess.py:71-74; torchaudio kaldi.py:436-511) */
int passt_mel_set_band(void* workspace, double fmin, double fmax, int sample_rate, void* stream);
/* same, (fmin, fmax) read from device memory (double[2]) when the kernel runs — identical launch every step, for CUDA graphs */
int passt_mel_set_band_dev(void* workspace, const double* band_dev, int sample_rate, void* stream);
/* wave f32 [B,L] -> log-mel f32 [B,128,T], T = 1+(L-1)/hop.  rnd: f32 [4,B] SpecAugment uniforms or NULL (eval);
 * pre-emphasis :59, STFT :60-61, power :62, mel matmul :76, log :78, freq/time masking :80-82, affine :84 */
int passt_mel_forward(const void* workspace, const float* wave, float* out, int B, int L, int hop, const float* rnd,
                      int freqm, int timem, void* stream);

/* ---- tcgen05 GEMM family: nn.Linear fwd/bwd (models/passt.py:279-289, :338-359) and PatchEmbed.proj (:315) ---- */
/* mode 0: C[M,N] = bf16(A[M,K] B[N,K]^T + bias)            (A,B,C bf16; K-major operands)
 * mode 1: C = bf16(gelu'(acc + bias)), C2 = bf16(gelu(acc + bias)) (Mlp.fc1 + nn.GELU, :285-286; exact erf)
 * mode 2: C f32 = acc + aux_f32[row % aux_period, :]        (patch-embed: bias + pos-embeds + cls/dist rows)
 * mode 3: C = bf16(acc * aux_bf16[row, :]) with aux = gelu'(pre) saved by mode 1 (fc2 dgrad fused with GELU
 *         backward); if `bias` is non-NULL it is an OUTPUT: bias[n] += sum_m C[m,n] (the fc1 bias gradient)
 * mode 4: C f32 [M,N] += A[K,M]^T B[K,N]                    (weight gradient; split-K, TMA reduce-add)
 * mode 0|16, 3|16: same as 0 / 3 with B given as [K,N] row-major — the nn.Linear weight itself, so input
 *         gradients need no transposed weight copy (the operand is read MN-major)                           */
int passt_gemm_bf16(const void* A, const void* B, void* C, void* C2, const float* bias, const void* aux, int M,
                    int N, int K, int lda, int ldb, int ldc, int mode, int aux_period, int ld_aux, int splits,
                    int max_ctas, void* stream);
/* 1 (default): 2-CTA tcgen05 (cta_group::2, 256x256 cluster tiles) where the shape allows; 0: 1-CTA kernel only */
void passt_gemm_set_2cta(int enable);
/* bring-up hook: override the UMMA shared-memory descriptor strides (6 uint32); active=0 restores defaults */
void passt_gemm_debug_desc(int active, const unsigned* v6);

/* ---- attention: softmax(q k^T * scale) v, models/passt.py:345-358, and its autograd ---------------------------- */
/* qkv bf16 [B,N,3*H*64] (layout of nn.Linear(dim,3*dim) output, :345) -> out bf16 [B,N,H*64],
 * lse f32 [B,H,Npad] (Npad = 128*ceil(N/128); log2-domain log-sum-exp of the scaled scores, pad rows +inf) */
int passt_attn_fwd(const void* qkv, void* out, float* lse, int B, int N, int H, float scale, void* stream);
/* bring-up hook: device buffer (>= 3*512 int64) receiving clock64 stamps of CTA 0 of attn_fwd; NULL disables */
void passt_attn_debug_timeline(void* buf);
void passt_attn_bwd_debug_timeline(void* buf);
size_t passt_attn_bwd_workspace_bytes(int B, int N, int H);
/* d_bias_qkv (optional, f32 [3*H*64]) += column sums of d_qkv: the bias gradient of the qkv Linear */
int passt_attn_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, void* d_qkv,
                   float* d_bias_qkv, void* workspace, int B, int N, int H, float scale, void* stream);

/* D-fusion: zero the workspace ahead of time, let the GEMM that produces d_out (passt_gemm_bf16 mode 5, kRowDotBf16)
 * accumulate D = rowsum(d_out o out) into passt_attn_bwd_dsum_ptr(workspace), then call passt_attn_bwd_ex with flags = 1
 * (no memset, no D pre-pass).  flags = 0 is passt_attn_bwd. */
int passt_attn_bwd_prepare(void* workspace, int B, int N, int H, void* stream);
float* passt_attn_bwd_dsum_ptr(void* workspace, int B, int N, int H);
int passt_attn_bwd_ex(const void* qkv, const void* out, const void* d_out, const float* lse, void* d_qkv,
                      float* d_bias_qkv, void* workspace, int B, int N, int H, float scale, int flags, void* stream);
/* forward schedule: 2 (default) = ping-pong kernel, one CTA per SM with two query tiles in flight (attn_fwd2.cu);
 * 4 = the same with the next Q K^T issued as soon as a tile's scores are in registers; 3 = ping-pong with two softmax
 * threads per row (attn_fwd3.cu); 1 = two CTAs per SM, one query tile each (attn_fwd.cu).  Same results. */
void passt_attn_fwd_set_variant(int variant);
/* attention backward kernel: 1 (default) = eight compute warps (64-query column halves), 2 = sixteen (32-query quarters);
 * environment PASST_B200_ATTN_BWD.  Same results. */
void passt_attn_bwd_set_variant(int variant);
int passt_attn_bwd_get_variant(void);
/* programmatic dependent launch of the hot-path kernels (1 = default; environment PASST_B200_PDL=0 turns it off) */
void passt_set_pdl(int enable);
int passt_get_pdl(void);
/* SMs the persistent kernels (GEMMs, attention) occupy, default 148: the data-parallel backward leaves a few SMs to
 * the NCCL all-reduce kernels that run concurrently (passt_b200/ddp.py) */
void passt_set_sm_limit(int n_sms);
int passt_get_sm_limit(void);

/* ---- row kernels ------------------------------------------------------------------------------------------------ */
/* x_out = x_in (+ delta); h = LayerNorm(x_out) (Block residual + norm1/norm2, models/passt.py:377-380) */
int passt_ln_fwd(const float* x_in, const void* delta_bf16, float* x_out, void* h_bf16, float* mean, float* rstd,
                 const float* gamma, const float* beta, int M, int dim, float eps, void* stream);
/* g_out = g_in + dLN(dh); bf16 copy; dgamma/dbeta/colsum accumulate (+=) */
int passt_ln_bwd(const void* dh_bf16, const float* x, const float* mean, const float* rstd, const float* gamma,
                 const float* g_in, float* g_out, void* g_out_bf16, float* dgamma, float* dbeta, float* colsum,
                 int M, int dim, void* stream);
/* out[n] += sum_m in[m,n]  (bias gradients) */
int passt_colsum_bf16(const void* in_bf16, float* out, int M, int N, int ld, void* stream);
/* kept 16x16 patches of mel f32 [B,Fm,Tm] -> bf16 rows [B*ntok,256]; rows 0,1 of each clip are zero (cls/dist);
 * optional fused spectrogram mixup (ex_audioset.py:173-177).  Patchout gathers of models/passt.py:535-552. */
int passt_im2col(const float* mel, void* A_bf16, const int* patch_f, const int* patch_t, int B, int ntok, int Fm,
                 int Tm, int fstride, int tstride, const int* mix_perm, const float* mix_lam, void* stream);
/* single-kernel patch embedding: TMA gather of the kept 16x16 patches (optionally mixing two clips) -> bf16 operand
 * tile in shared memory -> tcgen05 GEMM with the conv weight w_bf16 [768, 256] -> + token table -> f32 [B*ntok, 768]
 * (PatchEmbed.proj + positional adds + Patchout + token assembly, models/passt.py:315-323,527-564).  Returns a
 * negative code when Tm % 4 != 0 (TMA needs 16-byte aligned mel rows), Tm < 160 (one strip box) or ntok < 32: use
 * passt_im2col + passt_gemm_bf16 then. */
int passt_patch_embed(const float* mel, const void* w_bf16, const float* tab, float* out, const int* patch_f,
                      const int* patch_t, int B, int ntok, int Fm, int Tm, int fstride, int tstride,
                      const int* mix_perm, const float* mix_lam, void* stream);
/* additive token table: conv bias + time/freq pos-embed (models/passt.py:527-529), cls/dist rows (:557-564);
 * toff_dev (optional int*): time-embedding offset read from device memory instead of `toff` (CUDA graphs) */
int passt_token_table(float* tab, const float* cls, const float* dist, const float* new_pos,
                      const float* conv_bias, const float* time_pos, const float* freq_pos, const int* patch_f,
                      const int* patch_t, int ntok, int Fg, int Tg, int toff, const int* toff_dev, void* stream);
int passt_token_table_bwd(const float* g0, float* dcls, float* ddist, float* dnew_pos, float* dconv_bias,
                          float* dtime, float* dfreq, const int* patch_f, const int* patch_t, int B, int ntok,
                          int Fg, int Tg, int toff, const int* toff_dev, void* stream);
/* f32 [R,C] -> bf16 [R,C] and bf16 [C,R] (tensor-core operand copies of the fp32 master weights) */
int passt_cast_transpose(const float* in, void* out_bf16, void* outT_bf16, int R, int C, void* stream);

/* fp32 -> bf16 of a list of matrices in ONE launch (the per-step refresh of every bf16 weight copy; replaces the
 * reference's per-op autocast casts, torch/amp autocast of F.linear at models/passt.py:285-289,345,359).
 * table: device array of n_entries 32-byte records {const float* src; void* dst_bf16; uint64_t n8; uint32_t
 * first_block; uint32_t pad}, n8 = elements / 8, first_block = running sum of ceil(n8 / 1024); total_blocks = the
 * final sum. */
int passt_cast_multi(const void* table, int n_entries, int total_blocks, void* stream);

/* AdamW over a list of parameter tensors in ONE launch, refreshing the bf16 GEMM-operand copy of a weight in the same
 * pass (replaces torch.optim.AdamW as built by get_optimizer, ex_audioset.py:104-109, on the training hot path).
 * table: device array of n_entries 64-byte records {float* p; const float* g; float* m; float* v; void* w16_or_null;
 * uint64_t n; uint32_t first_block; uint32_t vec; uint64_t reserved}, first_block = running sum of ceil(n / 4096),
 * vec = 1 when n % 4 == 0 and all pointers are 16-byte aligned; total_blocks = the final sum.
 * hyper: device float[8] = {lr, beta1, beta2, eps, weight_decay, step, scratch, scratch}; step is advanced here. */
int passt_adamw_step(const void* table, int n_entries, int total_blocks, float* hyper, void* stream);
/* final norm on cls/dist rows, (cls+dist)/2, head LayerNorm + Linear (models/passt.py:570-588, :463-464) */
int passt_head_fwd(const float* x, const void* delta_bf16, const float* norm_g, const float* norm_b,
                   const float* hln_g, const float* hln_b, const float* W, const float* bias, float* logits,
                   float* features, float* fl, int B, int ntok, int C, void* stream);
int passt_head_bwd(const float* x, const void* delta_bf16, const float* norm_g, const float* norm_b,
                   const float* hln_g, const float* hln_b, const float* W, const float* dlogits,
                   const float* dfeatures, const float* fl, float* g_out, void* g_out_bf16, float* d_norm_g,
                   float* d_norm_b, float* d_hln_g, float* d_hln_b, float* dW, float* dbias, float* colsum, int B,
                   int ntok, int C, void* stream);


/* ---- losses, SWA, validation post-processing ------------------------------------- */
/* mean BCE-with-logits against targets mixed on the fly (y*lam + y[perm]*(1-lam); perm/lam NULL = no mixup) and
 * d loss / d logits in ONE launch (training_step, ex_audioset.py:172-192).  workspace: passt_loss_workspace_bytes(B)
 * bytes, zero-initialised once by the caller. */
size_t passt_loss_workspace_bytes(int B);
int passt_loss_bce(const float* logits, const float* target, const int* perm, const float* lam, float* loss,
                   float* dlogits, void* workspace, int B, int C, void* stream);
/* mean of CE(z, y)*lam + CE(z, y[perm])*(1-lam) and its logit gradient (ex_esc50.py:151-169); target: int64 [B] */
int passt_loss_ce(const float* logits, const long long* target, const int* perm, const float* lam, float* loss,
                  float* dlogits, void* workspace, int B, int C, void* stream);
/* out = in * scalar_dev[0]  (backward of the two losses: dlogits times the upstream gradient, no host sync) */
int passt_scale_dev(float* out, const float* in, const float* scalar_dev, size_t n, void* stream);
/* validation side: mean logits of K nets and sigmoid (mode 0: sigmoid(mean logits), EnsembelerModel models/passt.py:
 * 1021-1036 + torch.sigmoid of validation_step ex_audioset.py:236-238; mode 1: mean of sigmoids).  logits_ptrs is a
 * HOST array of K <= 16 device pointers. */
int passt_ens_sigmoid(const float* const* logits_ptrs, int K, float* mean_logits, float* prob, size_t n, int mode,
                      void* stream);
/* per-class average precision with sklearn's tie handling (validation_epoch_end, ex_audioset.py:262-266) on the
 * device: scores / targets f32 [n, C], ap f32 [C] (NaN for a class without positives); n <= ~45000 */
int passt_average_precision(const float* scores, const float* targets, float* ap, int n, int C, void* stream);
/* SWA running average over a list of tensors in ONE launch (helpers/swa_callback.py:246-268).  table: device array of
 * 32-byte records {const float* p_model; float* p_swa; uint64_t n; uint32_t first_block; uint32_t pad},
 * first_block = running sum of ceil(n / 4096); n_averaged == 0 copies. */
int passt_swa_update(const void* table, int n_entries, int total_blocks, long long n_averaged, void* stream);

/* ---- waveform-side augmentation while staging a batch ------------------------------ */
/* gain -> pad/truncate to L -> roll -> optional zero-mean waveform mixup with another clip of the batch (+ target
 * mix), audioset/dataset.py:107-140,315-339.  raw: all source clips (f32); src_off [B] int64 element offsets;
 * src_len [B]; gain [B] or NULL; shift [B] or NULL; mix_idx [B] (-1 = not mixed) with mix_lam [B], or both NULL;
 * out [B, L]; tgt / tgt_out [B, C] or both NULL. */
int passt_wave_augment(const float* raw, const long long* src_off, const int* src_len, const float* gain,
                       const int* shift, const int* mix_idx, const float* mix_lam, float* out, const float* tgt,
                       float* tgt_out, int B, int L, int C, void* stream);

/* ---- fp32-parity tier of the forward pass (north_star: 1e-3 vs the fp32 reference) ------------------------------- */
/* GEMM operands split into bf16 hi/lo parts so that the bf16 tcgen05 GEMM (passt_gemm_bf16, mode 2, fp32 output) over
 * a 3x longer contraction computes A_hi W_hi + A_hi W_lo + A_lo W_hi in fp32: fp32 [R, C] (row stride ld_in) ->
 * bf16 [R, 3C]; pattern 0 = [hi|hi|lo] (activations, F.linear inputs models/passt.py:285-289,345,359),
 * pattern 1 = [hi|lo|hi] (weights). */
int passt_split3_bf16(const float* in, void* out_bf16, long long R, int C, int ld_in, int pattern, void* stream);
/* split3(gelu_exact(in)) (models/passt.py:286-287) */
int passt_gelu_split3(const float* in, void* out_bf16, long long R, int C, void* stream);
/* residual add + LayerNorm in fp32 with the split operand of the next GEMM as output (models/passt.py:377-380) */
int passt_ln_fwd_f32tier(const float* x_in, const float* delta_f32, float* x_out, void* h_split_bf16,
                         const float* gamma, const float* beta, int M, int dim, float eps, void* stream);
/* ---- backward of the fp32 tier (gradients within 1e-3): split products with the contraction along ROWS ------------ */
/* fp32 [R, C] -> bf16 [3R, C]; pattern 0 = [hi; hi; lo] (A side), 1 = [hi; lo; hi] (B side).  dgrad dX = dY W uses
 * passt_split3_bf16(dY, pattern 0) x passt_split3_rows_bf16(W, pattern 1) through passt_gemm_bf16 mode 2|16 (fp32 out);
 * wgrad dW += dY^T X uses rows(dY, 0) x rows(X, 1) through mode 4. */
int passt_split3_rows_bf16(const float* in, void* out_bf16, long long R, int C, int ld_in, int pattern, void* stream);
/* bf16 [R, 3C] = [hi|hi|lo] (a forward operand) -> bf16 [3R, C] rows in pattern 0 / 1 */
int passt_restack3_bf16(const void* in_bf16, void* out_bf16, long long R, int C, int pattern, void* stream);
/* out[c] += sum_r in[r, c] in fp32 (bias gradients) */
int passt_colsum_f32(const float* in, float* out, long long R, int C, void* stream);
/* h = LayerNorm(x) * gamma + beta in fp32 (dim 768) */
int passt_ln_apply_f32(const float* x, float* h, const float* gamma, const float* beta, int M, int dim, float eps,
                       void* stream);
/* g_out = g_in (nullable) + dLN(dh; x, gamma); dgamma / dbeta += (all fp32; autograd of models/passt.py:377-380 norms) */
int passt_ln_bwd_f32(const float* dh, const float* x, const float* gamma, const float* g_in, float* g_out, float* dgamma,
                     float* dbeta, int M, int dim, float eps, void* stream);
/* dpre = dact * gelu'(pre), exact-erf GELU (models/passt.py:280) */
int passt_gelu_bwd_f32(const float* dact, const float* pre, float* dpre, long long n, void* stream);
/* attention backward in fp32: qkv f32 [B,N,3C], d_out f32 [B,N,C] -> d_qkv f32 [B,N,3C]; workspace 3*B*H*N floats
 * (autograd of models/passt.py:345-358) */
int passt_attn_bwd_f32(const float* qkv, const float* d_out, float* d_qkv, float* workspace, int B, int N, int H,
                       float scale, void* stream);
/* passt_im2col with fp32 output rows [B*ntok, 256] (split afterwards) */
int passt_im2col_f32(const float* mel, float* A_f32, const int* patch_f, const int* patch_t, int B, int ntok, int Fm,
                     int Tm, int fstride, int tstride, const int* mix_perm, const float* mix_lam, void* stream);
/* softmax(q k^T * scale) v in fp32 (models/passt.py:345-358): qkv f32 [B, N, 3*H*64] -> split3 of the [B, N, H*64]
 * attention output */
int passt_attn_fwd_f32(const float* qkv, void* out_split_bf16, int B, int N, int H, float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PASST_B200_H */
