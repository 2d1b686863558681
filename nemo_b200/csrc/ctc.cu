// CTC loss — the reference's `CTCLoss` (nemo/collections/asr/losses/ctc.py:25-82) is torch.nn.CTCLoss
// with blank = num_classes, reduction 'none' (then 'mean_batch' / 'mean_volume' on the host side, :57-60) applied to the
// log-probabilities of ConvASRDecoder.  This is the same alpha / beta recursion (Graves et al. 2006, eqs. 6-7, 10-11, 16)
// in log space, one CTA per utterance:
//   forward : log_alpha [B, T, S] (S = 2*Umax+1, kept for the backward), nll[b] = -logsumexp(alpha[T_b-1, S_b-1], alpha[T_b-1, S_b-2])
//   backward: log_beta kept in shared memory; dlogp[b,t,c] = g_b * (exp(lp[t,c]) - exp(logsum_{s: l'_s = c}(alpha+beta) + nll - lp[t,c]))
//             for t < T_b, 0 beyond — the formula torch's native kernel uses (it already contains the softmax Jacobian term).
// log_probs are read in the decoder's own [B, T, V+1] layout (the reference transposes to [T, B, V+1] first, ctc.py:75).
#include <math_constants.h>

#include "common.cuh"

namespace b2c {

constexpr int kCtcThreads = 256;

__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

__device__ __forceinline__ int ctc_label(const int64_t* tgt, int s, int blank) { return (s & 1) ? (int)tgt[s >> 1] : blank; }

__global__ void __launch_bounds__(kCtcThreads)
ctc_fwd_kernel(const float* __restrict__ logp, const int64_t* __restrict__ targets, const int64_t* __restrict__ in_len,
               const int64_t* __restrict__ tgt_len, float* __restrict__ alpha, float* __restrict__ nll, int T, int V1, int Umax,
               int blank) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, S = 2 * Umax + 1;
  const int Tb = (int)min((int64_t)T, in_len[b]), U = (int)min((int64_t)Umax, tgt_len[b]), Sb = 2 * U + 1;
  const int64_t* tgt = targets + (size_t)b * Umax;
  const float* lp = logp + (size_t)b * T * V1;
  float* al = alpha + (size_t)b * T * S;
  float* row0 = sm;
  float* row1 = sm + S;
  if (Tb <= 0) {
    if (threadIdx.x == 0) nll[b] = (U == 0) ? 0.f : CUDART_INF_F;
    return;
  }
  for (int s = threadIdx.x; s < S; s += kCtcThreads) {
    float v = -CUDART_INF_F;
    if (s == 0) v = lp[blank];
    else if (s == 1 && U > 0) v = lp[tgt[0]];
    row0[s] = v;
    al[s] = v;
  }
  __syncthreads();
  float* prev = row0;
  float* cur = row1;
  for (int t = 1; t < Tb; ++t) {
    const float* lpt = lp + (size_t)t * V1;
    for (int s = threadIdx.x; s < S; s += kCtcThreads) {
      float v = -CUDART_INF_F;
      if (s < Sb) {
        const int l = ctc_label(tgt, s, blank);
        const float a = prev[s];
        const float bb = s >= 1 ? prev[s - 1] : -CUDART_INF_F;
        const float c = ((s & 1) && s >= 3 && tgt[s >> 1] != tgt[(s >> 1) - 1]) ? prev[s - 2] : -CUDART_INF_F;
        v = lse3(a, bb, c) + lpt[l];
      }
      cur[s] = v;
      al[(size_t)t * S + s] = v;
    }
    __syncthreads();
    float* tmp = prev; prev = cur; cur = tmp;
  }
  if (threadIdx.x == 0) {
    const float a = prev[Sb - 1], bb = Sb >= 2 ? prev[Sb - 2] : -CUDART_INF_F;
    nll[b] = -lse3(a, bb, -CUDART_INF_F);
  }
}

__global__ void __launch_bounds__(kCtcThreads)
ctc_bwd_kernel(const float* __restrict__ logp, const int64_t* __restrict__ targets, const int64_t* __restrict__ in_len,
               const int64_t* __restrict__ tgt_len, const float* __restrict__ alpha, const float* __restrict__ nll,
               const float* __restrict__ grad_out, float* __restrict__ dlogp, int T, int V1, int Umax, int blank,
               int zero_infinity) {
  extern __shared__ float sm[];
  __shared__ float red[kCtcThreads / 32];
  __shared__ float bmax;
  const int b = blockIdx.x, S = 2 * Umax + 1;
  const int Tb = (int)min((int64_t)T, in_len[b]), U = (int)min((int64_t)Umax, tgt_len[b]), Sb = 2 * U + 1;
  const int64_t* tgt = targets + (size_t)b * Umax;
  const float* lp = logp + (size_t)b * T * V1;
  const float* al = alpha + (size_t)b * T * S;
  float* dl = dlogp + (size_t)b * T * V1;
  float* row0 = sm;            // beta ping
  float* row1 = sm + S;        // beta pong
  float* acc = sm + 2 * S;     // [V1] per-class sums of exp(alpha + beta - max)
  const float nl = nll[b];
  const float g = grad_out ? grad_out[b] : 1.f;
  const bool dead = zero_infinity && (nl == CUDART_INF_F);
  // frames beyond the input length, and everything of an infeasible utterance under zero_infinity: zero gradient
  for (size_t i = (size_t)(dead ? 0 : max(Tb, 0)) * V1 + threadIdx.x; i < (size_t)T * V1; i += kCtcThreads) dl[i] = 0.f;
  if (dead || Tb <= 0) return;
  float* nxt = row0;
  float* cur = row1;
  for (int t = Tb - 1; t >= 0; --t) {
    const float* lpt = lp + (size_t)t * V1;
    // ---- beta_t
    for (int s = threadIdx.x; s < S; s += kCtcThreads) {
      float v = -CUDART_INF_F;
      if (s < Sb) {
        const int l = ctc_label(tgt, s, blank);
        if (t == Tb - 1) {
          v = (s == Sb - 1 || s == Sb - 2) ? lpt[l] : -CUDART_INF_F;
        } else {
          const float a = nxt[s];
          const float bb = s + 1 < Sb ? nxt[s + 1] : -CUDART_INF_F;
          const float c = ((s & 1) && s + 2 < Sb && tgt[s >> 1] != tgt[(s >> 1) + 1]) ? nxt[s + 2] : -CUDART_INF_F;
          v = lse3(a, bb, c) + lpt[l];
        }
      }
      cur[s] = v;
    }
    for (int c = threadIdx.x; c < V1; c += kCtcThreads) acc[c] = 0.f;
    __syncthreads();
    // ---- per-class logsumexp of alpha + beta: block max, then shared-memory atomics of exp(. - max)
    float m = -CUDART_INF_F;
    for (int s = threadIdx.x; s < Sb; s += kCtcThreads) m = fmaxf(m, al[(size_t)t * S + s] + cur[s]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float mm = red[0];
      for (int w = 1; w < kCtcThreads / 32; ++w) mm = fmaxf(mm, red[w]);
      bmax = mm;
    }
    __syncthreads();
    const float mt = bmax;
    if (mt > -CUDART_INF_F)
      for (int s = threadIdx.x; s < Sb; s += kCtcThreads)
        atomicAdd(&acc[ctc_label(tgt, s, blank)], expf(al[(size_t)t * S + s] + cur[s] - mt));
    __syncthreads();
    for (int c = threadIdx.x; c < V1; c += kCtcThreads) {
      const float l = lpt[c];
      const float post = acc[c] > 0.f ? expf(logf(acc[c]) + mt + nl - l) : 0.f;
      dl[(size_t)t * V1 + c] = g * (expf(l) - post);
    }
    __syncthreads();
    float* tmp = nxt; nxt = cur; cur = tmp;
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" size_t b2c_ctc_workspace(int B, int T, int Umax) { return (size_t)B * T * (2 * (size_t)Umax + 1) * sizeof(float); }

extern "C" int b2c_ctc_loss_fwd(const float* logp, const int64_t* targets, const int64_t* in_len, const int64_t* tgt_len,
                                float* alpha_ws, float* nll, int B, int T, int V1, int Umax, int blank, void* stream) {
  B2C_REQUIRE(B > 0 && T > 0 && V1 > 1 && Umax >= 1, "ctc_loss_fwd: bad shape");
  B2C_REQUIRE(blank >= 0 && blank < V1, "ctc_loss_fwd: blank %d outside [0,%d)", blank, V1);
  const size_t smem = (size_t)2 * (2 * Umax + 1) * sizeof(float);
  B2C_REQUIRE(smem <= 200 * 1024, "ctc_loss_fwd: target length %d too long", Umax);
  static size_t set = 0;
  if (smem > 48 * 1024 && smem > set) {
    B2C_CUDA(cudaFuncSetAttribute(ctc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set = smem;
  }
  ctc_fwd_kernel<<<B, kCtcThreads, smem, (cudaStream_t)stream>>>(logp, targets, in_len, tgt_len, alpha_ws, nll, T, V1, Umax, blank);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_ctc_loss_bwd(const float* logp, const int64_t* targets, const int64_t* in_len, const int64_t* tgt_len,
                                const float* alpha_ws, const float* nll, const float* grad_out, float* dlogp, int B, int T,
                                int V1, int Umax, int blank, int zero_infinity, void* stream) {
  B2C_REQUIRE(B > 0 && T > 0 && V1 > 1 && Umax >= 1, "ctc_loss_bwd: bad shape");
  const size_t smem = ((size_t)2 * (2 * Umax + 1) + V1) * sizeof(float);
  B2C_REQUIRE(smem <= 200 * 1024, "ctc_loss_bwd: vocabulary %d / target length %d too large", V1, Umax);
  static size_t set = 0;
  if (smem > 48 * 1024 && smem > set) {
    B2C_CUDA(cudaFuncSetAttribute(ctc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set = smem;
  }
  ctc_bwd_kernel<<<B, kCtcThreads, smem, (cudaStream_t)stream>>>(logp, targets, in_len, tgt_len, alpha_ws, nll, grad_out, dlogp, T, V1,
                                                                 Umax, blank, zero_infinity);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
