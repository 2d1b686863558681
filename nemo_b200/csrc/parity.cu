// fp32-accurate "parity mode" of the forward pass (SURVEY.md §8c tolerance policy (i); north_star: encoder logits within
// 1e-3 relative of the reference).  The dense contractions still run on the tcgen05 GEMM core (csrc/gemm.cu) — every fp32
// operand x is split into two bf16 terms x = hi + lo (+ 2^-17 |x|), and A·B is accumulated as hi·hi + hi·lo + lo·hi in the
// fp32 TMEM / fp32-memory accumulator (the dropped lo·lo term is 2^-18 relative) — and everything between the GEMMs is
// plain fp32 here: no bf16 tensor is ever stored.  These kernels are deliberately simple (one thread per element): parity
// mode is for validating the speed-mode kernels against the fp32 oracle at the BASELINE shapes, not for throughput.
#include <math_constants.h>

#include "common.cuh"

namespace b2c {

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void p_split2_kernel(const float* __restrict__ x, bf16* __restrict__ hi, bf16* __restrict__ lo, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  const bf16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// out[m, n] = bias ? bias[n] : 0
__global__ void p_fill_rows_kernel(float* __restrict__ out, const float* __restrict__ bias, long long M, int N, long long ldo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const long long m = i / N;
  const int n = (int)(i - m * N);
  out[m * ldo + n] = bias ? bias[n] : 0.f;
}

enum { P_SILU = 0, P_AXPY = 1, P_GLU_MASK = 2, P_RELU_MASK = 3, P_ADD_ROWVEC = 4, P_DSILU = 5, P_GLU_BWD = 6, P_RELU_BWD = 7 };

__device__ __forceinline__ float dsilu_exact(float x) {
  const float s = sigmoid_exact(x);
  return s * (1.f + x * (1.f - s));
}

// elementwise ops on [M, N] fp32 (row pitch lda for `a`, N for the rest)
__global__ void p_pointwise_kernel(int mode, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                   long long M, int N, long long lda, float alpha, const int32_t* __restrict__ lengths,
                                   int rows_per_batch, int rows_inner) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const long long m = i / N;
  const int n = (int)(i - m * N);
  bool valid = true;
  if (lengths) {
    const long long bb = m / rows_per_batch;
    const long long t = (m - bb * rows_per_batch) / rows_inner;
    valid = t < lengths[bb];
  }
  float r;
  if (mode == P_SILU) {
    const float x = a[m * lda + n];
    r = x * sigmoid_exact(x);
  } else if (mode == P_AXPY) {  // out = a + alpha * b  (a may be null)
    r = (a ? a[m * lda + n] : 0.f) + alpha * b[i];
  } else if (mode == P_GLU_MASK) {  // a = [M, 2N] (value | gate), conformer_modules.py:325-331
    const float v = a[m * lda + n], g = a[m * lda + N + n];
    r = valid ? v * sigmoid_exact(g) : 0.f;
  } else if (mode == P_RELU_MASK) {  // subsampling.py:725-753: ReLU then time mask
    r = valid ? fmaxf(a[m * lda + n], 0.f) : 0.f;
  } else if (mode == P_ADD_ROWVEC) {  // out = a + b[n]   (q + pos_bias_u)
    r = a[m * lda + n] + b[n];
  } else if (mode == P_DSILU) {  // out = a * silu'(b)
    r = a[m * lda + n] * dsilu_exact(b[i]);
  } else if (mode == P_RELU_BWD) {  // out = b > 0 ? a : 0   (b = the forward's ReLU+mask output)
    r = b[i] > 0.f ? a[m * lda + n] : 0.f;
  } else {  // P_GLU_BWD: a = du [M,N], b = g [M,2N] (value | gate) → out [M,2N]
    const float du = valid ? a[m * lda + n] : 0.f;
    const float v = b[m * 2 * N + n], gt = b[m * 2 * N + N + n];
    const float sg = sigmoid_exact(gt);
    out[m * 2 * N + n] = du * sg;
    out[m * 2 * N + N + n] = du * v * sg * (1.f - sg);
    return;
  }
  out[i] = r;
}

// out[n] (+)= sum_m x[m, n]   (double accumulation; x has row pitch ld)
__global__ void __launch_bounds__(256) p_colsum_kernel(const float* __restrict__ x, long long ld, float* __restrict__ out, long long M,
                                                       int N, int accumulate) {
  __shared__ double s1[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), r0 = threadIdx.x >> 5;
  double a = 0.0;
  if (c < N)
    for (long long m = r0; m < M; m += 8) a += (double)x[m * ld + c];
  s1[r0][threadIdx.x & 31] = a;
  __syncthreads();
  if (r0 == 0 && c < N) {
    for (int k = 1; k < 8; ++k) a += s1[k][threadIdx.x];
    out[c] = (accumulate ? out[c] : 0.f) + (float)a;
  }
}

// BatchNorm + SiLU backward, elementwise part: dz = ds * silu'(c*scale+shift); dzx = dz * xhat  (xhat = (c-mean)*rstd)
__global__ void p_bn_silu_bwd_pre_kernel(const float* __restrict__ ds, const float* __restrict__ c, const float* __restrict__ scale,
                                         const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd,
                                         float* __restrict__ dz, float* __restrict__ dzx, long long M, int N) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int n = (int)(i % N);
  const float v = ds[i] * dsilu_exact(c[i] * scale[n] + shift[n]);
  dz[i] = v;
  dzx[i] = v * (c[i] - mean[n]) * rstd[n];
}
// dc = gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat)) (train) | gamma*rstd*dz (eval)
__global__ void p_bn_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ c, const float* __restrict__ gamma,
                                      const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ sum_dz,
                                      const float* __restrict__ sum_dzx, float* __restrict__ dc, long long M, int N, int training) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int n = (int)(i % N);
  const float gr = gamma[n] * rstd[n];
  float v = dz[i];
  if (training) {
    const float inv = 1.f / (float)M;
    v = v - sum_dz[n] * inv - (c[i] - mean[n]) * rstd[n] * sum_dzx[n] * inv;
  }
  dc[i] = gr * v;
}

// du[b,t,c] = (t < len_b) * sum_j w[c,j] * dc[b, t + pad - j, c]
__global__ void p_dwconv1d_bwd_data_kernel(const float* __restrict__ dc, const float* __restrict__ w, const int32_t* __restrict__ lengths,
                                           float* __restrict__ du, int B, int T, int d, int k) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T * d) return;
  const int c = (int)(i % d);
  const int t = (int)((i / d) % T);
  const int b = (int)(i / ((long long)d * T));
  const int pad = (k - 1) / 2;
  float acc = 0.f;
  for (int j = 0; j < k; ++j) {
    const int tt = t + pad - j;
    if (tt >= 0 && tt < T) acc = fmaf(w[c * k + j], dc[((size_t)b * T + tt) * d + c], acc);
  }
  du[i] = (!lengths || t < lengths[b]) ? acc : 0.f;
}
// dw[c,j] += sum_{b,t} dc[b,t,c] * u[b,t+j-pad,c]
__global__ void p_dwconv1d_bwd_weight_kernel(const float* __restrict__ dc, const float* __restrict__ u, float* __restrict__ dw, int B,
                                             int T, int d, int k) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // j * d + c  (c fastest: coalesced)
  if (idx >= d * k) return;
  const int c = idx % d, j = idx / d, pad = (k - 1) / 2;
  double acc = 0.0;
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T; ++t) {
      const int tt = t + j - pad;
      if (tt >= 0 && tt < T) acc += (double)dc[((size_t)b * T + t) * d + c] * (double)u[((size_t)b * T + tt) * d + c];
    }
  dw[c * k + j] += (float)acc;
}

// softmax backward + un-skew: rowdot = sum_j P dP; dS = P (dP - rowdot) scale; dR[i, T-1-i+j] = dS[i,j], 0 elsewhere
__global__ void __launch_bounds__(256) p_attn_softmax_bwd_kernel(const float* __restrict__ P, const float* __restrict__ dP,
                                                                 float* __restrict__ dS, float* __restrict__ dR, long long rows, int T,
                                                                 int Tp, int Rp, float scale) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int i = (int)(row % T);
  const float* p = P + row * Tp;
  const float* dp = dP + row * Tp;
  float dot = 0.f;
  for (int j = lane; j < T; j += 32) dot = fmaf(p[j], dp[j], dot);
  dot = warp_sum(dot);
  float* ds = dS + row * Tp;
  float* dr = dR + row * Rp;
  for (int x = lane; x < Rp; x += 32) dr[x] = 0.f;
  __syncwarp();
  for (int j = lane; j < Tp; j += 32) {
    float v = 0.f;
    if (j < T) {
      v = p[j] * (dp[j] - dot) * scale;
      dr[T - 1 - i + j] = v;
    }
    ds[j] = v;
  }
}

// transpose of p_im2col_nhwc as a gather (no atomics): dx[b,t1,f1,ci] = sum over the (t2,f2,tap) that read it of dp[row, tap*C+ci]
__global__ void p_col2im_nhwc_kernel(const float* __restrict__ dp, float* __restrict__ dx, int B, int T1, int F1, int C, int T2, int F2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T1 * F1 * C) return;
  const int ci = (int)(i % C);
  const int f1 = (int)((i / C) % F1);
  const int t1 = (int)((i / ((long long)C * F1)) % T1);
  const int b = (int)(i / ((long long)C * F1 * T1));
  float acc = 0.f;
  for (int kh = 0; kh < 3; ++kh) {
    const int tn = t1 + 1 - kh;
    if (tn < 0 || (tn & 1) || (tn >> 1) >= T2) continue;
    for (int kw = 0; kw < 3; ++kw) {
      const int fn = f1 + 1 - kw;
      if (fn < 0 || (fn & 1) || (fn >> 1) >= F2) continue;
      const long long row = ((long long)b * T2 + (tn >> 1)) * F2 + (fn >> 1);
      acc += dp[row * 9 * C + (kh * 3 + kw) * C + ci];
    }
  }
  dx[i] = acc;
}

// depthwise 3x3 s2 backward (dy rows t2 >= len_out count as zero): data gradient (gather) and weight / bias gradient
__global__ void p_dwconv2d_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w, const int32_t* __restrict__ len_out,
                                           float* __restrict__ dx, int B, int T1, int F1, int C, int T2, int F2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T1 * F1 * C) return;
  const int c = (int)(i % C);
  const int f1 = (int)((i / C) % F1);
  const int t1 = (int)((i / ((long long)C * F1)) % T1);
  const int b = (int)(i / ((long long)C * F1 * T1));
  float acc = 0.f;
  for (int kh = 0; kh < 3; ++kh) {
    const int tn = t1 + 1 - kh;
    if (tn < 0 || (tn & 1) || (tn >> 1) >= T2 || (tn >> 1) >= len_out[b]) continue;
    for (int kw = 0; kw < 3; ++kw) {
      const int fn = f1 + 1 - kw;
      if (fn < 0 || (fn & 1) || (fn >> 1) >= F2) continue;
      acc = fmaf(w[c * 9 + kh * 3 + kw], dy[(((size_t)b * T2 + (tn >> 1)) * F2 + (fn >> 1)) * C + c], acc);
    }
  }
  dx[i] = acc;
}
__global__ void p_dwconv2d_bwd_weight_kernel(const float* __restrict__ dy, const float* __restrict__ x, const int32_t* __restrict__ len_out,
                                             float* __restrict__ dw, float* __restrict__ dbias, int B, int T1, int F1, int C, int T2,
                                             int F2) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // tap * C + c, tap 9 = bias
  if (idx >= 10 * C) return;
  const int c = idx % C, tap = idx / C;
  double acc = 0.0;
  for (int b = 0; b < B; ++b) {
    const int lo = min(len_out[b], T2);
    for (int t2 = 0; t2 < lo; ++t2)
      for (int f2 = 0; f2 < F2; ++f2) {
        const double g = (double)dy[(((size_t)b * T2 + t2) * F2 + f2) * C + c];
        if (tap == 9) { acc += g; continue; }
        const int t = 2 * t2 + tap / 3 - 1, f = 2 * f2 + tap % 3 - 1;
        if (t >= 0 && t < T1 && f >= 0 && f < F1) acc += g * (double)x[(((size_t)b * T1 + t) * F1 + f) * C + c];
      }
  }
  if (tap == 9) dbias[c] += (float)acc;
  else dw[c * 9 + tap] += (float)acc;
}

// y = silu(c * scale[n] + shift[n])
__global__ void p_bn_silu_kernel(const float* __restrict__ c, const float* __restrict__ scale, const float* __restrict__ shift,
                                 float* __restrict__ out, long long M, int N) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int n = (int)(i % N);
  const float x = c[i] * scale[n] + shift[n];
  out[i] = x * sigmoid_exact(x);
}

// per-column sum and sum of squares of [M, N] (double accumulation), one block of 256 threads per 32 columns
__global__ void __launch_bounds__(256) p_colstats_kernel(const float* __restrict__ x, float* __restrict__ sum, float* __restrict__ sumsq,
                                                         long long M, int N) {
  __shared__ double s1[8][33], s2[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), r0 = threadIdx.x >> 5;
  double a = 0.0, q = 0.0;
  if (c < N)
    for (long long m = r0; m < M; m += 8) {
      const double v = (double)x[m * N + c];
      a += v;
      q += v * v;
    }
  s1[r0][threadIdx.x & 31] = a;
  s2[r0][threadIdx.x & 31] = q;
  __syncthreads();
  if (r0 == 0 && c < N) {
    for (int k = 1; k < 8; ++k) { a += s1[k][threadIdx.x]; q += s2[k][threadIdx.x]; }
    sum[c] = (float)a;
    sumsq[c] = (float)q;
  }
}

// Conv2d(1->C) stem patches from the mel layout [B, F, T]: out [B*T1*F1, 16] fp32, columns kh*3+kw (kh = time tap), zero padded;
// input frames >= len_in are zero (MaskedConvSequential masks the input first, subsampling.py:733-735)
__global__ void p_im2col_mel_kernel(const float* __restrict__ mel, const int32_t* __restrict__ len_in, float* __restrict__ out,
                                    int B, int F, int T, int T1, int F1) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long rows = (long long)B * T1 * F1;
  if (i >= rows * 16) return;
  const long long row = i >> 4;
  const int col = (int)(i & 15);
  float v = 0.f;
  if (col < 9) {
    const int f1 = (int)(row % F1);
    const int t1 = (int)((row / F1) % T1);
    const int b = (int)(row / ((long long)F1 * T1));
    const int t = 2 * t1 + col / 3 - 1, f = 2 * f1 + col % 3 - 1;
    if (t >= 0 && t < T && f >= 0 && f < F && t < len_in[b]) v = mel[((size_t)b * F + f) * T + t];
  }
  out[i] = v;
}

// 3x3 stride-2 pad-1 patches of a channels-last fp32 tensor [B, T1, F1, C]: out [B*T2*F2, 9*C], k = tap*C + ci (tap = kh*3+kw)
__global__ void p_im2col_nhwc_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int T1, int F1, int C, int T2,
                                     int F2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long K = 9LL * C, rows = (long long)B * T2 * F2;
  if (i >= rows * K) return;
  const long long row = i / K;
  const int k = (int)(i - row * K), tap = k / C, ci = k - tap * C;
  const int f2 = (int)(row % F2);
  const int t2 = (int)((row / F2) % T2);
  const int b = (int)(row / ((long long)F2 * T2));
  const int t = 2 * t2 + tap / 3 - 1, f = 2 * f2 + tap % 3 - 1;
  out[i] = (t >= 0 && t < T1 && f >= 0 && f < F1) ? x[(((size_t)b * T1 + t) * F1 + f) * C + ci] : 0.f;
}

// depthwise Conv2d(C, 3x3, s2, p1, groups=C) channels-last + bias, rows t2 >= len_out zeroed (subsampling.py:142-215)
__global__ void p_dwconv2d_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                  const int32_t* __restrict__ len_out, float* __restrict__ out, int B, int T1, int F1, int C, int T2,
                                  int F2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T2 * F2 * C) return;
  const int c = (int)(i % C);
  const int f2 = (int)((i / C) % F2);
  const int t2 = (int)((i / ((long long)C * F2)) % T2);
  const int b = (int)(i / ((long long)C * F2 * T2));
  float acc = bias[c];
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      const int t = 2 * t2 + kh - 1, f = 2 * f2 + kw - 1;
      if (t >= 0 && t < T1 && f >= 0 && f < F1) acc = fmaf(w[c * 9 + kh * 3 + kw], x[(((size_t)b * T1 + t) * F1 + f) * C + c], acc);
    }
  out[i] = (t2 < len_out[b]) ? acc : 0.f;
}

// depthwise k-tap Conv1d over time, symmetric padding, channels-last [B, T, d] (causal_convs.py:130-150)
__global__ void p_dwconv1d_kernel(const float* __restrict__ u, const float* __restrict__ w, const float* __restrict__ bias,
                                  float* __restrict__ out, int B, int T, int d, int k) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * T * d) return;
  const int c = (int)(i % d);
  const int t = (int)((i / d) % T);
  const int b = (int)(i / ((long long)d * T));
  const int pad = (k - 1) / 2;
  float acc = bias[c];
  for (int j = 0; j < k; ++j) {
    const int tt = t + j - pad;
    if (tt >= 0 && tt < T) acc = fmaf(w[c * k + j], u[((size_t)b * T + tt) * d + c], acc);
  }
  out[i] = acc;
}

// attention probabilities of one (b, h, i) row per warp:
//   s[j] = scale * (AC[i, j] + R[i, T-1-i+j]);  masked_fill(-1e4) ; softmax ; masked_fill(0)   (multi_head_attention.py:125-146,
//   259-270, 344-349).  AC [BH, T, Tp], R [BH, T, Rp], P [BH, T, Tp] (pad columns of P are zeroed).
__global__ void __launch_bounds__(256) p_attn_softmax_kernel(const float* __restrict__ AC, const float* __restrict__ R,
                                                             const int32_t* __restrict__ lengths, float* __restrict__ P, int B, int H,
                                                             int T, int Tp, int Rp, float scale) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * H * T) return;
  const int i = (int)(row % T);
  const int b = (int)(row / ((long long)T * H));
  const int len = lengths ? min(lengths[b], T) : T;
  const float* ac = AC + row * Tp;
  const float* r = R + row * Rp + (T - 1 - i);
  float* p = P + row * Tp;
  const bool qvalid = i < len;
  float mx = -CUDART_INF_F;
  for (int j = lane; j < T; j += 32) {
    const float s = (qvalid && j < len) ? scale * (ac[j] + r[j]) : -10000.f;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < T; j += 32) {
    const float s = (qvalid && j < len) ? scale * (ac[j] + r[j]) : -10000.f;
    sum += expf(s - mx);
  }
  sum = warp_sum(sum);
  for (int j = lane; j < Tp; j += 32) {
    float v = 0.f;
    if (j < T && qvalid && j < len) v = expf(scale * (ac[j] + r[j]) - mx) / sum;
    p[j] = v;
  }
}

// log_softmax over the V classes of fp32 logits [M, ld] → [M, V] (conv_asr.py:456-468), one warp per row
__global__ void __launch_bounds__(256) p_log_softmax_kernel(const float* __restrict__ logits, int ld, float* __restrict__ out, long long M,
                                                            int V) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* x = logits + row * ld;
  float mx = -CUDART_INF_F;
  for (int j = lane; j < V; j += 32) mx = fmaxf(mx, x[j]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < V; j += 32) s += expf(x[j] - mx);
  const float lse = mx + logf(warp_sum(s));
  for (int j = lane; j < V; j += 32) out[row * V + j] = x[j] - lse;
}

static inline unsigned nblk(long long n, int t = 256) { return (unsigned)((n + t - 1) / t); }

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_p_split2(const float* x, void* hi_bf16, void* lo_bf16, long long n, void* stream) {
  if (n <= 0) return B2C_OK;
  p_split2_kernel<<<nblk(n), 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<bf16*>(hi_bf16), reinterpret_cast<bf16*>(lo_bf16), n);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_fill_rows(float* out, const float* bias, long long M, int N, long long ldo, void* stream) {
  p_fill_rows_kernel<<<nblk(M * N), 256, 0, (cudaStream_t)stream>>>(out, bias, M, N, ldo);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_pointwise(int mode, const float* a, const float* b, float* out, long long M, int N, long long lda, float alpha,
                               const int32_t* lengths, int rows_per_batch, int rows_inner, void* stream) {
  B2C_REQUIRE(mode >= 0 && mode <= 7, "p_pointwise: mode %d", mode);
  p_pointwise_kernel<<<nblk(M * N), 256, 0, (cudaStream_t)stream>>>(mode, a, b, out, M, N, lda, alpha, lengths,
                                                                    rows_per_batch > 0 ? rows_per_batch : 1, rows_inner > 0 ? rows_inner : 1);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_bn_silu(const float* c, const float* scale, const float* shift, float* out, long long M, int N, void* stream) {
  p_bn_silu_kernel<<<nblk(M * N), 256, 0, (cudaStream_t)stream>>>(c, scale, shift, out, M, N);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_colstats(const float* x, float* sum, float* sumsq, long long M, int N, void* stream) {
  p_colstats_kernel<<<(N + 31) / 32, 256, 0, (cudaStream_t)stream>>>(x, sum, sumsq, M, N);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_im2col_mel(const float* mel, const int32_t* len_in, float* out, int B, int F, int T, void* stream) {
  const int T1 = (T - 1) / 2 + 1, F1 = (F - 1) / 2 + 1;
  p_im2col_mel_kernel<<<nblk((long long)B * T1 * F1 * 16), 256, 0, (cudaStream_t)stream>>>(mel, len_in, out, B, F, T, T1, F1);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_im2col_nhwc(const float* x, float* out, int B, int T1, int F1, int C, void* stream) {
  const int T2 = (T1 - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  p_im2col_nhwc_kernel<<<nblk((long long)B * T2 * F2 * 9 * C), 256, 0, (cudaStream_t)stream>>>(x, out, B, T1, F1, C, T2, F2);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_dwconv2d(const float* x, const float* w, const float* bias, const int32_t* len_out, float* out, int B, int T1,
                              int F1, int C, void* stream) {
  const int T2 = (T1 - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  p_dwconv2d_kernel<<<nblk((long long)B * T2 * F2 * C), 256, 0, (cudaStream_t)stream>>>(x, w, bias, len_out, out, B, T1, F1, C, T2, F2);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_dwconv1d(const float* u, const float* w, const float* bias, float* out, int B, int T, int d, int k, void* stream) {
  p_dwconv1d_kernel<<<nblk((long long)B * T * d), 256, 0, (cudaStream_t)stream>>>(u, w, bias, out, B, T, d, k);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_attn_softmax(const float* AC, const float* R, const int32_t* lengths, float* P, int B, int H, int T, int Tp, int Rp,
                                  float scale, void* stream) {
  B2C_REQUIRE(Tp >= T && Rp >= 2 * T - 1, "p_attn_softmax: pitches too small");
  p_attn_softmax_kernel<<<nblk((long long)B * H * T * 32), 256, 0, (cudaStream_t)stream>>>(AC, R, lengths, P, B, H, T, Tp, Rp, scale);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_log_softmax(const float* logits, int ld, float* out, long long M, int V, void* stream) {
  p_log_softmax_kernel<<<nblk(M * 32), 256, 0, (cudaStream_t)stream>>>(logits, ld, out, M, V);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_colsum(const float* x, long long ld, float* out, long long M, int N, int accumulate, void* stream) {
  p_colsum_kernel<<<(N + 31) / 32, 256, 0, (cudaStream_t)stream>>>(x, ld, out, M, N, accumulate);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_bn_silu_bwd(const float* ds, const float* c, const float* gamma, const float* scale, const float* shift,
                                 const float* mean, const float* rstd, float* dz_scratch, float* dzx_scratch, float* sums, float* dc,
                                 long long M, int N, int training, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  p_bn_silu_bwd_pre_kernel<<<nblk(M * N), 256, 0, st>>>(ds, c, scale, shift, mean, rstd, dz_scratch, dzx_scratch, M, N);
  B2C_LAUNCH_CHECK();
  p_colsum_kernel<<<(N + 31) / 32, 256, 0, st>>>(dz_scratch, N, sums, M, N, 0);       // sums[0:N]  = sum dz   (= dbeta)
  B2C_LAUNCH_CHECK();
  p_colsum_kernel<<<(N + 31) / 32, 256, 0, st>>>(dzx_scratch, N, sums + N, M, N, 0);  // sums[N:2N] = sum dz*xhat (= dgamma)
  B2C_LAUNCH_CHECK();
  p_bn_bwd_apply_kernel<<<nblk(M * N), 256, 0, st>>>(dz_scratch, c, gamma, mean, rstd, sums, sums + N, dc, M, N, training);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_dwconv1d_bwd(const float* dc, const float* u, const float* w, const int32_t* lengths, float* du, float* dw, int B,
                                  int T, int d, int k, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  p_dwconv1d_bwd_data_kernel<<<nblk((long long)B * T * d), 256, 0, st>>>(dc, w, lengths, du, B, T, d, k);
  B2C_LAUNCH_CHECK();
  p_dwconv1d_bwd_weight_kernel<<<nblk((long long)d * k, 128), 128, 0, st>>>(dc, u, dw, B, T, d, k);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_attn_softmax_bwd(const float* P, const float* dP, float* dS, float* dR, int B, int H, int T, int Tp, int Rp,
                                      float scale, void* stream) {
  const long long rows = (long long)B * H * T;
  p_attn_softmax_bwd_kernel<<<nblk(rows * 32), 256, 0, (cudaStream_t)stream>>>(P, dP, dS, dR, rows, T, Tp, Rp, scale);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_col2im_nhwc(const float* dp, float* dx, int B, int T1, int F1, int C, void* stream) {
  const int T2 = (T1 - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  p_col2im_nhwc_kernel<<<nblk((long long)B * T1 * F1 * C), 256, 0, (cudaStream_t)stream>>>(dp, dx, B, T1, F1, C, T2, F2);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
extern "C" int b2c_p_dwconv2d_bwd(const float* dy, const float* x, const float* w, const int32_t* len_out, float* dx, float* dw,
                                  float* dbias, int B, int T1, int F1, int C, void* stream) {
  const int T2 = (T1 - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  cudaStream_t st = (cudaStream_t)stream;
  p_dwconv2d_bwd_data_kernel<<<nblk((long long)B * T1 * F1 * C), 256, 0, st>>>(dy, w, len_out, dx, B, T1, F1, C, T2, F2);
  B2C_LAUNCH_CHECK();
  p_dwconv2d_bwd_weight_kernel<<<nblk(10LL * C, 128), 128, 0, st>>>(dy, x, len_out, dw, dbias, B, T1, F1, C, T2, F2);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
