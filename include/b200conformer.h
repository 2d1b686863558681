/* libb200conformer — C ABI of the B200-native Conformer-ASR encoder hot path.
 *
 * The reference (NVIDIA-NeMo/NeMo) has no C/FFI plugin interface for this path: its boundary is the Python
 * module contract of `AudioToMelSpectrogramPreprocessor.forward` (nemo/collections/asr/modules/audio_preprocessing.py:93-103)
 * and `ConformerEncoder.forward` (nemo/collections/asr/modules/conformer_encoder.py:548-591), whose arithmetic
 * executes in torch ops (cuFFT / cuDNN / cuBLAS / ATen).  This header is the thin C ABI *under* that Python
 * boundary: plain device pointers + sizes, caller-owned memory, an explicit CUDA stream, no torch
 * types, no allocation, no implicit synchronisation.  `nemo_b200/` binds it with ctypes; INTEGRATION.md shows the
 * binding a NeMo maintainer would add.
 *
 * Every function returns B2C_OK (0) or a negative error code; `b2c_last_error()` returns a thread-local message.
 * All pointers are DEVICE pointers unless stated otherwise.  `stream` is a cudaStream_t passed as void*.
 */
#ifndef B200CONFORMER_H_
#define B200CONFORMER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2C_OK 0
#define B2C_ERR_INVALID (-1)
#define B2C_ERR_CUDA (-2)
#define B2C_ERR_UNSUPPORTED (-3)

const char* b2c_last_error(void);
int b2c_version(void);
/* number of kernels this library has launched since load (bench.py's `gpu_launches`). */
long long b2c_launch_count(void);

/* ------------------------------------------------------------------------------------------------ GEMM core
 * D[M,N] = sum_k A(m,k) B(n,k), bf16 operands, fp32 accumulation (tcgen05 / TMEM), fused epilogue.
 * a_mn_major = 0: A is [M, K] row-major (K contiguous, pitch lda elements); 1: A is [K, M] row-major.
 * b_mn_major likewise for B ([N, K] vs [K, N]).
 * Replaces nn.Linear / 1x1 Conv1d forward and autograd on the hot path (conformer_modules.py:322,347,382-387;
 * multi_head_attention.py:104-123,146,300; subsampling.py:431). */
typedef enum {
  B2C_EPI_BF16 = 0,      /* out0(bf16)[m,n] = acc + bias[n]                                               */
  B2C_EPI_SILU = 1,      /* out0 = z = bf16(acc+bias); out1 = bf16(dropout(silu(acc+bias)))  (FFN linear1)  */
  B2C_EPI_RESID_F32 = 2, /* out0(f32) = (aux? aux(f32) : 0) + alpha * dropout(bf16(acc+bias))  (residual)  */
  B2C_EPI_GLU_MASK = 3,  /* interleaved (value,gate) columns: out0 = bf16(acc+bias) [M,N];
                            out1[m,n/2] = bf16(a*sigmoid(g)) * (t<len)                (pointwise_conv1+GLU) */
  B2C_EPI_DSILU = 4,     /* out0 = bf16(dropout(acc) * silu'(aux(bf16)[m,n]))              (FFN backward)  */
  B2C_EPI_RELU_MASK = 5, /* out0 = bf16(relu(acc+bias)) * (t<len)                      (subsampling conv) */
  B2C_EPI_ACCUM_F32 = 6, /* out0(f32)[m,n] += alpha*acc  (split-K, atomic)             (weight gradients)  */
  B2C_EPI_RELU_BWD = 7   /* out0 = bf16(aux(bf16)[m,n] > 0 ? acc : 0)          (gradient through ReLU+mask) */
} b2c_epilogue_mode;

typedef struct {
  void* out0;
  void* out1;
  const void* aux;
  const float* bias;      /* [N] fp32 or NULL */
  const int32_t* lengths; /* [B] valid rows per batch item (GLU/RELU masks) or NULL */
  long long ldo;          /* pitch (elements) of out0 / aux */
  int rows_per_batch;     /* rows of the GEMM per batch item (T, or T*rows_inner) */
  int rows_inner;         /* RELU_MASK: rows per time step (F) */
  float alpha;
  unsigned long long seed; /* dropout: key of the stateless counter hash (element kept iff hash(seed, site, m*N+n) < keep) */
  unsigned int drop_site;  /* dropout: site id (layer*16 + site) */
  unsigned int drop_keep;  /* dropout: 16-bit keep threshold = round((1-p)*65536), 0 = no dropout; needs N % 16 == 0 */
  float drop_scale;        /* 1/(1-p) */
} b2c_epilogue;

int b2c_gemm(const void* A, const void* B, int M, int N, int K, int lda, int ldb, int a_mn_major, int b_mn_major,
             int epilogue_mode, const b2c_epilogue* epi, void* stream);

/* ------------------------------------------------------------------------------------------------ mel front-end
 * Fused FilterbankFeatures.forward (nemo/collections/asr/parts/preprocessing/features.py:423-502), eval/no-dither
 * or dither via `dither`>0 with Philox(seed): pre-emphasis, length mask, centred 512-pt STFT (hann 400, hop 160),
 * power, Slaney mel (sparse), log(x+2^-24), per-feature normalisation (unbiased std + 1e-5), frame masking.
 *   x [B,S] f32, length [B] i64  →  mel [B,n_mels,T] f32 (T = S/hop+1), seq_len [B] i64.
 * `fb` is the dense [n_mels, n_fft/2+1] f32 filterbank buffer (state_dict `featurizer.fb`), `window` [win] f32,
 * `fb_ranges` [n_mels*2] i32 = per-filter [first, last+1) non-zero bin, prepared once by b2c_mel_fb_ranges. */
int b2c_mel_seq_len(const int64_t* length, int64_t* seq_len, int B, int n_fft, int hop, void* stream);
int b2c_mel_fb_ranges(const float* fb, int n_mels, int n_bins, int32_t* ranges, void* stream);
int b2c_mel_frontend(const float* x, const int64_t* length, const float* window, const float* fb,
                     const int32_t* fb_ranges, float* mel, int64_t* seq_len, int B, int S, int T, int n_mels, int win,
                     int hop, int n_fft, float preemph, float dither, unsigned long long seed, int normalize,
                     float pad_value, void* stream);

/* ------------------------------------------------------------------------------------------------ elementwise / norm
 * LayerNorm over the last dim (eps 1e-5), fp32 in → bf16 out (+ optional fp32 out), saves mean / rstd per row.
 * torch.nn.LayerNorm under bf16 autocast (conformer_modules.py:94-158). */
int b2c_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, float* mean,
                      float* rstd, int M, int d, void* stream);
/* dy (bf16 if dy_is_bf16 else f32) → dx: dr[m,:] (+)= LN'(dy); dgamma/dbeta += ; optionally emits
 * next_bf16 = bf16(next_scale * dr) (the gradient entering the previous sub-block's output GEMM). */
int b2c_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* gamma, const float* mean,
                      const float* rstd, float* dr, int accumulate, float* dgamma, float* dbeta, void* next_bf16,
                      float next_scale, float drop_p, unsigned long long seed, unsigned int site, int M, int d,
                      void* stream);
/* out_f32[n] += sum_m in_bf16[m,n]  (bias gradients) */
int b2c_colsum_bf16(const void* in_bf16, float* out, int M, int N, int ld, void* stream);
/* fp32 → bf16 cast with optional scale */
int b2c_cast_bf16(const float* in, void* out_bf16, long long n, float scale, void* stream);
/* out_bf16[i] = bf16(scale * dropout_site(in[i]))  (backward of xscale + pre-encoder dropout) */
int b2c_scale_drop_cast(const float* in, void* out_bf16, long long n, float scale, float drop_p,
                        unsigned long long seed, unsigned int site, void* stream);
/* in [A,B,C] f32 → out permuted; mode 0: out[b][a][c]; mode 1: out[a][c][b].  Weight layout preparation
 * (GLU value/gate interleave, conv taps-major, flatten order) and the inverse for their gradients. */
int b2c_permute3_bf16(const float* in, void* out_bf16, int A, int B, int C, int mode, void* stream);
int b2c_permute3_f32(const float* in, float* out, int A, int B, int C, int mode, int accumulate, void* stream);
int b2c_fill_f32(float* p, long long n, float v, void* stream);

/* CTC head (ConvASRDecoder.forward, nemo/collections/asr/modules/conv_asr.py:456-468): the 1x1 Conv1d is a b2c_gemm;
 * these are the row log-softmax over the V classes of bf16 logits [M, ld] (ld >= V) → fp32 log-probs [M, V], and its
 * backward dlogits = dlogp - exp(logp) * rowsum(dlogp) (bf16 [M, ld], pad columns zeroed). */
int b2c_log_softmax_fwd(const void* logits_bf16, int ld, float* out, int M, int V, void* stream);
int b2c_log_softmax_bwd(const float* dlogp, const float* logp, void* dlogits_bf16, int ld, int M, int V, void* stream);

/* ------------------------------------------------------------------------------------------------ conv module
 * Depthwise k-tap conv over time (symmetric pad (k-1)/2) + bias on channels-last [B*T, d] bf16 with per-channel
 * sum / sum-of-squares accumulation for BatchNorm (causal_convs.py:130-150, conformer_modules.py:332-342). */
int b2c_dwconv_fwd(const void* u_bf16, const float* w, const float* bias, void* c_bf16, float* bn_sum, float* bn_sumsq,
                   int B, int T, int d, int k, void* stream);
/* BatchNorm1d finalise: train → batch mean/var (biased) into scale/shift and momentum update of running stats
 * (unbiased var), `num_batches_tracked` (int64 device word, may be NULL) += 1; eval → running stats.
 * scale = gamma*rstd, shift = beta - mean*scale. */
int b2c_bn_finalize(const float* bn_sum, const float* bn_sumsq, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float* scale, float* shift, float* mean_out,
                    float* rstd_out, long long count, int d, int training, float momentum, float eps,
                    long long* num_batches_tracked, void* stream);
/* s = bf16(silu(bf16(c*scale+shift))) */
int b2c_bn_silu_fwd(const void* c_bf16, const float* scale, const float* shift, void* s_bf16, long long M, int d,
                    void* stream);
/* BN(train)+SiLU backward, phase 1: per-channel sums of dz and dz*xhat (dz = ds*silu'(bn) ) */
int b2c_bn_silu_bwd_reduce(const void* ds_bf16, const void* c_bf16, const float* scale, const float* shift,
                           const float* mean, const float* rstd, float* sum_dz, float* sum_dz_xhat, long long M, int d,
                           void* stream);
/* phase 2: dc = gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat)) (train) or gamma*rstd*dz (eval); also dgamma/dbeta */
int b2c_bn_silu_bwd_apply(const void* ds_bf16, const void* c_bf16, const float* gamma, const float* scale,
                          const float* shift, const float* mean, const float* rstd, const float* sum_dz,
                          const float* sum_dz_xhat, void* dc_bf16, float* dgamma, float* dbeta, long long M, int d,
                          int training, void* stream);
/* depthwise conv backward: du (masked by t<len), dw [d,k], dbias [d] */
int b2c_dwconv_bwd(const void* dc_bf16, const void* u_bf16, const float* w, const int32_t* lengths, void* du_bf16,
                   float* dw, float* dbias, int B, int T, int d, int k, void* stream);
/* GLU backward on interleaved (value,gate) layout: dg[m,2j] = du*sig(g); dg[m,2j+1] = du*a*sig(g)(1-sig(g)) */
int b2c_glu_bwd(const void* du_bf16, const void* g_bf16, void* dg_bf16, long long M, int d, void* stream);

/* ------------------------------------------------------------------------------------------------ attention
 * Transformer-XL relative-position attention, flash-style (nothing [T,T]-shaped touches HBM):
 *   S[i,j] = ((q_i+u).k_j + (q_i+v).p[T-1-i+j]) / sqrt(dk);  masked_fill(-1e4); softmax; masked_fill(0); dropout; PV
 * (multi_head_attention.py:125-146,259-354).  qkv [B*T, 3*d] bf16 (q | k | v), pos [2T-1, d] bf16,
 * pos_bias_u/v [h, dk] f32, lengths [B] i32.  ctx [B*T, d] bf16, lse [B, h, T] f32 (base-2 log-sum-exp).
 * ctx_f32 (optional, [B*T, d] f32): un-rounded context kept for the backward's row term delta = sum(dO * O);
 * with only the bf16 context the rounding error biases whole rows of dS when value rows are nearly identical. */
int b2c_relpos_attn_fwd(const void* qkv, const void* pos, const float* bias_u, const float* bias_v,
                        const int32_t* lengths, void* ctx, float* ctx_f32, float* lse, int B, int T, int h, int dk,
                        float drop_p, unsigned long long seed, unsigned int site, void* stream);
size_t b2c_relpos_attn_bwd_workspace(int B, int T, int h);
int b2c_relpos_attn_bwd(const void* qkv, const void* pos, const float* bias_u, const float* bias_v,
                        const int32_t* lengths, const void* ctx, const float* ctx_f32, const void* dctx, const float* lse,
                        void* dqkv, float* dpos_f32, float* dbias_u, float* dbias_v, void* workspace, size_t workspace_bytes,
                        int B, int T, int h, int dk, float drop_p, unsigned long long seed, unsigned int site,
                        void* stream);

/* SpecAugment applied in one pass (spectr_augment.py:135-215, SURVEY §8f "next"): out[b,f,t] = mask_value inside any of
 * the utterance's n_t time spans [t_start, t_end) or n_f frequency bands [f_start, f_end) (arrays [B,n_t] / [B,n_f] i32,
 * drawn by the host module with the reference's own random draws), else in[b,f,t].  in may equal out. */
int b2c_spec_augment(const float* in, float* out, const int32_t* t_start, const int32_t* t_end, const int32_t* f_start,
                     const int32_t* f_end, int B, int F, int T, int n_t, int n_f, float mask_value, void* stream);

/* ------------------------------------------------------------------------------------------------ subsampling
 * First Conv2d(1→C, 3x3, stride 2, pad 1) + bias + ReLU + time masks, channels-last output [B, T1, F1, C] bf16
 * (subsampling.py:217-253,725-753).  mel [B, F, T] f32 (the encoder's input layout), w [C,1,3,3] f32,
 * len_in / len_out [B] i32 (valid frames before / after this conv). */
int b2c_subsample_conv1_fwd(const float* mel, const float* w, const float* bias, const int32_t* len_in,
                            const int32_t* len_out, void* out_bf16, int B, int F, int T, int C, void* stream);
/* dw [C,9] += , dbias [C] +=  from dout [B,T1,F1,C] bf16 (already ReLU/mask-gated) */
int b2c_subsample_conv1_bwd(const float* mel, const void* dout_bf16, const int32_t* len_in, float* dw, float* dbias,
                            int B, int F, int T, int C, void* stream);
/* Tensor-core route for the same stem: im2col of the masked mel input into patches [B*T1*F1, 16] bf16 (columns 0..8 =
 * the 3x3 window at kh*3+kw, column 9 = 1, 10..15 = 0).  Forward = b2c_gemm(patches, w16 [C,16], RELU_MASK epilogue);
 * weight + bias gradient = b2c_gemm(dout^T, patches, ACCUM_F32) → [C,16] whose column 9 is the bias gradient. */
int b2c_subsample_im2col(const float* mel, const int32_t* len_in, void* patches_bf16, int B, int F, int T, void* stream);
/* Conv2d(C→C, 3x3, stride 2, pad 1) as an implicit GEMM on tcgen05 (9 shifted, strided 4-D TMA tile loads feed one
 * TMEM accumulator; no im2col): in [B,T1,F1,C] bf16 → out [B,T2,F2,C] bf16 = relu(conv+bias) * (t2 < len_out).
 * w_bf16 [C_out, 9*C_in] with k = (kh*3+kw)*C_in + ci. */
int b2c_conv3x3s2_fwd(const void* in_bf16, const void* w_bf16, const float* bias, const int32_t* len_out,
                      void* out_bf16, int B, int T1, int F1, int C, void* stream);
/* data gradient: dx [B,T1,F1,C] = relu'(x) * conv_transpose(dy [B,T2,F2,C]); wd_bf16 [C_in, 9*C_out] with
 * k = (kh*3+kw)*C_out + co.  Four parity-class implicit GEMMs. */
int b2c_conv3x3s2_bwd_data(const void* dy_bf16, const void* wd_bf16, const void* x_bf16, void* dx_bf16, int B, int T1,
                           int F1, int C, void* stream);
/* weight gradient: dw f32 [C_out, 9*C_in] += (tokens are the contraction dimension; both operands gathered by TMA) */
int b2c_conv3x3s2_bwd_weight(const void* dy_bf16, const void* x_bf16, float* dw, int B, int T1, int F1, int C,
                             void* stream);
/* depthwise Conv2d(C, 3x3, s2, p1, groups=C) channels-last + bias, output rows t2 >= len_out zeroed (dw_striding,
 * subsampling.py:142-215). */
int b2c_dwconv3x3s2_fwd(const void* in_bf16, const float* w, const float* bias, const int32_t* len_out, void* out_bf16,
                        int B, int T1, int F1, int C, void* stream);

/* backward of the depthwise stage: dx = relu'(x) * conv_transpose(dy) (dy rows t2 >= len_out are treated as zero),
 * dw [C,9] +=, dbias [C] += */
int b2c_dwconv3x3s2_bwd(const void* dy_bf16, const void* x_bf16, const float* w, const int32_t* len_out, void* dx_bf16,
                        float* dw, float* dbias, int B, int T1, int F1, int C, void* stream);

/* ------------------------------------------------------------------------------------------------ misc
 * relative positional table rows = positions T-1 … -(T-1) (multi_head_attention.py:1007-1020,1056-1100) */
int b2c_relpos_table(void* pos_bf16, float* pos_f32, int T, int d, void* stream);
/* lengths after n stride-2 3x3 pad-1 convs: (L+2-3)//2+1 (subsampling.py:576-586,720-722).
 * len_in32 [B] = int32 copy of the input, len_stage [n_convs*B] int32 per stage, len_out64 [B]. */
int b2c_subsample_lengths(const int64_t* len_in, int32_t* len_in32, int32_t* len_stage, int64_t* len_out64, int B,
                          int n_convs, void* stream);

/* ------------------------------------------------------------------------------------------------ fp32 parity mode
 * Forward pass with fp32-accurate arithmetic (SURVEY.md §8c policy (i)): contractions stay on b2c_gemm — each fp32 operand is
 * split into bf16 hi + lo (b2c_p_split2) and A·B accumulates hi·hi + hi·lo + lo·hi through B2C_EPI_ACCUM_F32 into an fp32
 * output pre-filled with the bias (b2c_p_fill_rows) — and the ops between them are the plain fp32 kernels below (fp32 in,
 * fp32 out, exact expf sigmoid).  `nemo_b200.B200ConformerEncoder(precision="fp32")` orchestrates them. */
int b2c_p_split2(const float* x, void* hi_bf16, void* lo_bf16, long long n, void* stream);
int b2c_p_fill_rows(float* out, const float* bias, long long M, int N, long long ldo, void* stream);
/* mode 0: out = silu(a); 1: out = a + alpha*b (a may be NULL); 2: GLU of a [M,2N] (value | gate) * (t < len);
 * 3: relu(a) * (t < len); 4: out = a + b[n] (row vector).  `a` has row pitch lda; t = (m % rows_per_batch) / rows_inner. */
int b2c_p_pointwise(int mode, const float* a, const float* b, float* out, long long M, int N, long long lda, float alpha,
                    const int32_t* lengths, int rows_per_batch, int rows_inner, void* stream);
int b2c_p_bn_silu(const float* c, const float* scale, const float* shift, float* out, long long M, int N, void* stream);
int b2c_p_colstats(const float* x, float* sum, float* sumsq, long long M, int N, void* stream);
int b2c_p_im2col_mel(const float* mel, const int32_t* len_in, float* out, int B, int F, int T, void* stream);
int b2c_p_im2col_nhwc(const float* x, float* out, int B, int T1, int F1, int C, void* stream);
int b2c_p_dwconv2d(const float* x, const float* w, const float* bias, const int32_t* len_out, float* out, int B, int T1, int F1,
                   int C, void* stream);
int b2c_p_dwconv1d(const float* u, const float* w, const float* bias, float* out, int B, int T, int d, int k, void* stream);
/* P[bh,i,:] = masked softmax_j(scale * (AC[bh,i,j] + R[bh,i,T-1-i+j])) with the reference's -1e4 / 0 double mask */
int b2c_p_log_softmax(const float* logits, int ld, float* out, long long M, int V, void* stream);
/* backward of the above (same conventions; pointwise modes 5: out = a * silu'(b); 6: GLU backward, a = du [M,N], b = g [M,2N] →
 * out [M,2N]; 7: out = b > 0 ? a : 0).  out[n] (+)= column sums; BatchNorm+SiLU backward (sums [2N] = dbeta | dgamma on return);
 * depthwise conv data / weight gradients; softmax backward + un-skew (dS [BH,T,Tp], dR [BH,T,Rp]); col2im as a gather. */
int b2c_p_colsum(const float* x, long long ld, float* out, long long M, int N, int accumulate, void* stream);
int b2c_p_bn_silu_bwd(const float* ds, const float* c, const float* gamma, const float* scale, const float* shift, const float* mean,
                      const float* rstd, float* dz_scratch, float* dzx_scratch, float* sums, float* dc, long long M, int N,
                      int training, void* stream);
int b2c_p_dwconv1d_bwd(const float* dc, const float* u, const float* w, const int32_t* lengths, float* du, float* dw, int B, int T,
                       int d, int k, void* stream);
int b2c_p_attn_softmax_bwd(const float* P, const float* dP, float* dS, float* dR, int B, int H, int T, int Tp, int Rp, float scale,
                           void* stream);
int b2c_p_col2im_nhwc(const float* dp, float* dx, int B, int T1, int F1, int C, void* stream);
int b2c_p_dwconv2d_bwd(const float* dy, const float* x, const float* w, const int32_t* len_out, float* dx, float* dw, float* dbias,
                       int B, int T1, int F1, int C, void* stream);
int b2c_p_attn_softmax(const float* AC, const float* R, const int32_t* lengths, float* P, int B, int H, int T, int Tp, int Rp,
                       float scale, void* stream);

/* CTC loss (nemo/collections/asr/losses/ctc.py:25-82 = torch.nn.CTCLoss(blank=num_classes, reduction='none')): log-space
 * alpha / beta recursions, one CTA per utterance.  logp [B,T,V1] f32 (ConvASRDecoder layout), targets [B,Umax] i64 (padded),
 * in_len / tgt_len [B] i64.  Forward keeps log_alpha [B,T,2*Umax+1] in `alpha_ws` (b2c_ctc_workspace bytes) and writes the
 * per-utterance negative log-likelihood; backward writes dlogp [B,T,V1] = grad_out[b] * (exp(lp) - posterior), zero beyond
 * in_len (and everywhere for infeasible utterances when zero_infinity). */
size_t b2c_ctc_workspace(int B, int T, int Umax);
int b2c_ctc_loss_fwd(const float* logp, const int64_t* targets, const int64_t* in_len, const int64_t* tgt_len, float* alpha_ws,
                     float* nll, int B, int T, int V1, int Umax, int blank, void* stream);
int b2c_ctc_loss_bwd(const float* logp, const int64_t* targets, const int64_t* in_len, const int64_t* tgt_len,
                     const float* alpha_ws, const float* nll, const float* grad_out, float* dlogp, int B, int T, int V1, int Umax,
                     int blank, int zero_infinity, void* stream);

/* ------------------------------------------------------------------------------------------------ training-step runtime
 * Device-side RNG counter: every dropout / dither key becomes `seed + *counter * golden-ratio`, so a CUDA-graph replay of a
 * captured step (whose by-value seeds are frozen) still draws fresh masks each step.  b2c_set_seed_counter registers the
 * device word (NULL = by-value seeds only, the default); b2c_counter_add advances it (first node of a captured step). */
int b2c_set_seed_counter(const unsigned long long* counter);
int b2c_counter_add(unsigned long long* counter, unsigned long long inc, void* stream);

/* Multi-tensor tables (device arrays).  `blocks` is an int32 [n_blocks, 2] map block -> (tensor index, 2048-element chunk). */
typedef struct {
  void* ptr;          /* fp32 parameter tensor (contiguous) */
  long long offset;   /* its offset (elements) in the flat gradient / moment buffers */
  long long numel;
  int flags;          /* bit 0: no weight decay for this tensor */
  int reserved;
} b2c_tensor_entry;

typedef struct {
  float lr, beta1, beta2, eps, weight_decay;
  int noam_d_model;   /* > 0: NoamAnnealing (lr_scheduler.py:518-575): lr * d_model^-0.5 * min(s^-0.5, s * warmup^-1.5) */
  int noam_warmup;
  float noam_min_lr;
} b2c_adamw_hyper;

/* One optimizer step over all tensors of `table`: AdamW with decoupled weight decay (torch.optim.AdamW arithmetic =
 * NeMo's `optim.name: adamw`, nemo/core/classes/modelPT.py:650-823; conformer_ctc_bpe.yaml:175-192) reading the flat fp32
 * gradient buffer (scaled by grad_scale), exp_avg / exp_avg_sq flat fp32 moments.  `state` [1] f64 = step count (device,
 * incremented here), `scalars` [4] f32 scratch.  The learning rate is evaluated on the device: capturable in a CUDA graph.
 * zero_grad != 0: every consumed gradient element is overwritten with 0 (the next step's `zero_grad()` for free: the buffer is
 * written while it is being read, instead of a separate fill pass over it). */
int b2c_adamw_step(const b2c_tensor_entry* table, const int32_t* blocks, int n_blocks, float* grad_flat, float* exp_avg,
                   float* exp_avg_sq, double* state, float* scalars, const b2c_adamw_hyper* hyper, float grad_scale,
                   int zero_grad, void* stream);

typedef struct {
  const void* src;    /* fp32 [A, B, C] contiguous */
  void* dst;          /* bf16 (or fp32 if flags & 2): element (a, b, c) is written at a*sa + b*sb + c*sc */
  int A, B, C;
  int flags;
  long long sa, sb, sc;
} b2c_prep_entry;

/* fp32 master weights -> operand copies in kernel layouts (casts, GLU interleave, conv taps-major, flatten order), one
 * launch for the whole encoder; replaces per-tensor b2c_cast_bf16 / b2c_permute3_* calls of the forward. */
int b2c_prep_weights(const b2c_prep_entry* table, const int32_t* blocks, int n_blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200CONFORMER_H_ */
