// K2: first subsampling convolution Conv2d(1→C, 3x3, stride 2, pad 1) + bias + ReLU + time masks, forward and
// weight-gradient, on CUDA cores (C_in = 1: a 9-tap stencil, HBM-bound on its [B,T1,F1,C] bf16 output — no GEMM).
// Reference: ConvSubsampling 'striding' / 'dw_striding' layer 0 + MaskedConvSequential
// (nemo/collections/asr/parts/submodules/subsampling.py:142-253,725-753) under bf16 autocast
// (inputs and weights rounded to bf16, fp32 accumulate, bf16 output).
// Also the depthwise 3x3 stride-2 stages of dw_striding (channels-last, TMA-free register stencil).
#include "common.cuh"

namespace b2c {

// one CTA per (b, t1): thread = channel pair, loop over f1.  mel [B, F, T] fp32 → out [B, T1, F1, C] bf16
__global__ void __launch_bounds__(256)
subsample_conv1_fwd_kernel(const float* __restrict__ mel, const float* __restrict__ w, const float* __restrict__ bias,
                           const int32_t* __restrict__ len_in, const int32_t* __restrict__ len_out,
                           bf16* __restrict__ out, int F, int T, int C, int T1, int F1) {
  extern __shared__ float sx[];  // [3][F + 2]
  const int t1 = blockIdx.x, b = blockIdx.y;
  const int lin = len_in[b], lout = len_out[b];
  const int FP = F + 2;
  for (int i = threadIdx.x; i < 3 * FP; i += blockDim.x) {
    const int kh = i / FP, fp = i - kh * FP;
    const int t = 2 * t1 - 1 + kh, f = fp - 1;
    float v = 0.f;
    if (t >= 0 && t < T && t < lin && f >= 0 && f < F) v = bf16_round(mel[((size_t)b * F + f) * T + t]);
    sx[i] = v;
  }
  __syncthreads();
  const bool valid = t1 < lout;
  for (int cp = threadIdx.x; cp * 2 < C; cp += blockDim.x) {
    const int c = cp * 2;
    float w0[9], w1[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      w0[j] = bf16_round(w[(size_t)c * 9 + j]);
      w1[j] = bf16_round(w[(size_t)(c + 1) * 9 + j]);
    }
    const float b0 = bias[c], b1 = bias[c + 1];
    bf16* orow = out + (((size_t)b * T1 + t1) * F1) * C + c;
    for (int f1 = 0; f1 < F1; ++f1) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float x = sx[kh * FP + 2 * f1 + kw];
          a0 = fmaf(w0[kh * 3 + kw], x, a0);
          a1 = fmaf(w1[kh * 3 + kw], x, a1);
        }
      // conv output (bf16) → mask → ReLU → mask  (subsampling.py:731-752)
      const float r0 = valid ? fmaxf(bf16_round(a0 + b0), 0.f) : 0.f;
      const float r1 = valid ? fmaxf(bf16_round(a1 + b1), 0.f) : 0.f;
      *reinterpret_cast<uint32_t*>(orow + (size_t)f1 * C) = pack_bf16(r0, r1);
    }
  }
}

// im2col of the 1-channel stem for the tensor-core path: patches [B*T1*F1, 16] bf16, row (b,t1,f1) = the nine inputs
// mask(mel)[b, 2f1-1+kw, 2t1-1+kh] at column kh*3+kw, 1.0 at column 9 (so the weight-gradient GEMM's column 9 is the
// bias gradient), zeros in 10..15.  One CTA = 32 output time steps of one utterance; the mel tile is read once, coalesced.
constexpr int kImT = 32;
__global__ void __launch_bounds__(256)
subsample_im2col_kernel(const float* __restrict__ mel, const int32_t* __restrict__ len_in, bf16* __restrict__ patches, int F,
                        int T, int T1, int F1) {
  extern __shared__ float tile[];  // [F][2*kImT + 3]
  constexpr int W = 2 * kImT + 3;
  const int b = blockIdx.y, t1_0 = blockIdx.x * kImT;
  const int lin = len_in[b];
  for (int i = threadIdx.x; i < F * W; i += blockDim.x) {
    const int f = i / W, tt = i - f * W;
    const int t = 2 * t1_0 - 1 + tt;
    float v = 0.f;
    if (tt < 2 * kImT + 1 && t >= 0 && t < T && t < lin) v = mel[((size_t)b * F + f) * T + t];
    tile[i] = v;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < kImT * F1; r += blockDim.x) {
    const int tl = r / F1, f1 = r - tl * F1;
    const int t1 = t1_0 + tl;
    if (t1 >= T1) break;
    float x[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int f = 2 * f1 - 1 + kw;
        x[kh * 3 + kw] = (f >= 0 && f < F) ? tile[f * W + 2 * tl + kh] : 0.f;
      }
    uint4* dst = reinterpret_cast<uint4*>(patches + (((size_t)b * T1 + t1) * F1 + f1) * 16);
    dst[0] = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
    dst[1] = make_uint4(pack_bf16(x[8], 1.f), 0u, 0u, 0u);
  }
}

// dW[c, kh, kw] += sum dY[b,t1,f1,c] * x[b, 2t1+kh-1, 2f1+kw-1];  db[c] += sum dY   (dY already ReLU/mask gated)
__global__ void __launch_bounds__(256)
subsample_conv1_bwd_kernel(const float* __restrict__ mel, const bf16* __restrict__ dy, const int32_t* __restrict__ len_in,
                           float* __restrict__ dw, float* __restrict__ dbias, int F, int T, int C, int T1, int F1,
                           int rows_per_cta) {
  extern __shared__ float sx[];  // [3][F + 2]
  const int b = blockIdx.y;
  const int lin = len_in[b];
  const int FP = F + 2;
  const int t1_0 = blockIdx.x * rows_per_cta, t1_1 = min(T1, t1_0 + rows_per_cta);
  // this implementation keeps one channel pair per thread: C <= 2 * blockDim.x * passes
  for (int cp0 = 0; cp0 * 2 < C; cp0 += blockDim.x) {
    const int cp = cp0 + threadIdx.x;
    const bool act = cp * 2 < C;
    const int c = cp * 2;
    float g0[9], g1[9], d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) g0[j] = g1[j] = 0.f;
    for (int t1 = t1_0; t1 < t1_1; ++t1) {
      __syncthreads();
      for (int i = threadIdx.x; i < 3 * FP; i += blockDim.x) {
        const int kh = i / FP, fp = i - kh * FP;
        const int t = 2 * t1 - 1 + kh, f = fp - 1;
        float v = 0.f;
        if (t >= 0 && t < T && t < lin && f >= 0 && f < F) v = bf16_round(mel[((size_t)b * F + f) * T + t]);
        sx[i] = v;
      }
      __syncthreads();
      if (act) {
        const bf16* drow = dy + (((size_t)b * T1 + t1) * F1) * C + c;
        for (int f1 = 0; f1 < F1; ++f1) {
          const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(drow + (size_t)f1 * C));
          d0 += g.x; d1 += g.y;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const float x = sx[kh * FP + 2 * f1 + kw];
              g0[kh * 3 + kw] = fmaf(g.x, x, g0[kh * 3 + kw]);
              g1[kh * 3 + kw] = fmaf(g.y, x, g1[kh * 3 + kw]);
            }
        }
      }
    }
    if (act) {
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        atomicAdd(dw + (size_t)c * 9 + j, g0[j]);
        atomicAdd(dw + (size_t)(c + 1) * 9 + j, g1[j]);
      }
      atomicAdd(dbias + c, d0);
      atomicAdd(dbias + c + 1, d1);
    }
  }
}

// depthwise Conv2d(C, 3x3, s2, p1, groups=C) on channels-last [B,T1,F1,C] → [B,T2,F2,C], + bias, output masked t2<len
// (input is already time-masked by its producer).  thread = channel pair; CTA = (b, t2) row.
__global__ void __launch_bounds__(256)
dwconv3x3s2_fwd_kernel(const bf16* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                       const int32_t* __restrict__ len_out, bf16* __restrict__ out, int T1, int F1, int C, int T2, int F2) {
  const int t2 = blockIdx.x, b = blockIdx.y;
  const bool valid = len_out ? (t2 < len_out[b]) : true;
  for (int cp = threadIdx.x; cp * 2 < C; cp += blockDim.x) {
    const int c = cp * 2;
    float w0[9], w1[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      w0[j] = bf16_round(w[(size_t)c * 9 + j]);
      w1[j] = bf16_round(w[(size_t)(c + 1) * 9 + j]);
    }
    const float b0 = bias[c], b1 = bias[c + 1];
    for (int f2 = 0; f2 < F2; ++f2) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int t = 2 * t2 - 1 + kh;
        if (t < 0 || t >= T1) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int f = 2 * f2 - 1 + kw;
          if (f < 0 || f >= F1) continue;
          const float2 x = unpack_bf16(*reinterpret_cast<const uint32_t*>(in + (((size_t)b * T1 + t) * F1 + f) * C + c));
          a0 = fmaf(w0[kh * 3 + kw], x.x, a0);
          a1 = fmaf(w1[kh * 3 + kw], x.y, a1);
        }
      }
      const uint32_t o = valid ? pack_bf16(a0 + b0, a1 + b1) : 0u;
      *reinterpret_cast<uint32_t*>(out + (((size_t)b * T2 + t2) * F2 + f2) * C + c) = o;
    }
  }
}

// depthwise 3x3 s2 backward, data: dx[b,t1,f1,c] = (x[b,t1,f1,c] > 0) * sum_{kh,kw} dy[b,(t1+1-kh)/2,(f1+1-kw)/2,c] * w[c,kh,kw]
// (dy rows t2 >= len_out count as zero: the forward masks the depthwise output before the pointwise conv)
__global__ void __launch_bounds__(256)
dwconv3x3s2_bwd_data_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ w,
                            const int32_t* __restrict__ len_out, bf16* __restrict__ dx, int T1, int F1, int C, int T2,
                            int F2) {
  const int t1 = blockIdx.x, b = blockIdx.y;
  const int lout = len_out ? len_out[b] : T2;
  for (int cp = threadIdx.x; cp * 2 < C; cp += blockDim.x) {
    const int c = cp * 2;
    float w0[9], w1[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      w0[j] = bf16_round(w[(size_t)c * 9 + j]);
      w1[j] = bf16_round(w[(size_t)(c + 1) * 9 + j]);
    }
    for (int f1 = 0; f1 < F1; ++f1) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int tn = t1 + 1 - kh;
        if (tn < 0 || (tn & 1)) continue;
        const int t2 = tn >> 1;
        if (t2 >= T2 || t2 >= lout) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int fn = f1 + 1 - kw;
          if (fn < 0 || (fn & 1)) continue;
          const int f2 = fn >> 1;
          if (f2 >= F2) continue;
          const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(dy + (((size_t)b * T2 + t2) * F2 + f2) * C + c));
          a0 = fmaf(w0[kh * 3 + kw], g.x, a0);
          a1 = fmaf(w1[kh * 3 + kw], g.y, a1);
        }
      }
      const size_t o = (((size_t)b * T1 + t1) * F1 + f1) * C + c;
      const float2 xv = unpack_bf16(*reinterpret_cast<const uint32_t*>(x + o));
      *reinterpret_cast<uint32_t*>(dx + o) = pack_bf16(xv.x > 0.f ? a0 : 0.f, xv.y > 0.f ? a1 : 0.f);
    }
  }
}

// weight / bias gradients of the depthwise 3x3 s2 conv: dw[c,kh,kw] += sum dy[b,t2,f2,c] * x[b,2t2+kh-1,2f2+kw-1,c]
__global__ void __launch_bounds__(256)
dwconv3x3s2_bwd_weight_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const int32_t* __restrict__ len_out,
                              float* __restrict__ dw, float* __restrict__ dbias, int T1, int F1, int C, int T2, int F2,
                              int rows_per_cta) {
  const int b = blockIdx.y;
  const int lout = len_out ? min(len_out[b], T2) : T2;
  const int t2_0 = blockIdx.x * rows_per_cta, t2_1 = min(lout, t2_0 + rows_per_cta);
  for (int cp = threadIdx.x; cp * 2 < C; cp += blockDim.x) {
    const int c = cp * 2;
    float g0[9], g1[9], d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) g0[j] = g1[j] = 0.f;
    for (int t2 = t2_0; t2 < t2_1; ++t2) {
      for (int f2 = 0; f2 < F2; ++f2) {
        const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(dy + (((size_t)b * T2 + t2) * F2 + f2) * C + c));
        d0 += g.x; d1 += g.y;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int t = 2 * t2 - 1 + kh;
          if (t < 0 || t >= T1) continue;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const int f = 2 * f2 - 1 + kw;
            if (f < 0 || f >= F1) continue;
            const float2 xv = unpack_bf16(*reinterpret_cast<const uint32_t*>(x + (((size_t)b * T1 + t) * F1 + f) * C + c));
            g0[kh * 3 + kw] = fmaf(g.x, xv.x, g0[kh * 3 + kw]);
            g1[kh * 3 + kw] = fmaf(g.y, xv.y, g1[kh * 3 + kw]);
          }
        }
      }
    }
    if (t2_1 > t2_0) {
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        atomicAdd(dw + (size_t)c * 9 + j, g0[j]);
        atomicAdd(dw + (size_t)(c + 1) * 9 + j, g1[j]);
      }
      atomicAdd(dbias + c, d0);
      atomicAdd(dbias + c + 1, d1);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 8-channel (16-byte) variants of the three depthwise 3x3 stride-2 kernels.  The 2-channel versions above keep only 4-byte
// loads in flight (measured: ~12 % issue utilisation, 28 % of the HBM roofline on [16,1501,40,256]); here a thread owns 8
// consecutive channels of one output position, issues all of a position's 16-byte tap loads before the first FMA, and reads
// its 72 weights from a shared-memory copy [tap][C] made once per CTA.  thread = (channel group cg = tid % (C/8), position
// lane tid / (C/8)); a warp's 32 lanes read 512 contiguous bytes of one (t, f) position.
constexpr int kDw8Threads = 256;

__device__ __forceinline__ void unpack8f(const uint4& u, float (&v)[8]) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void load_dw_weights(float* sw, const float* __restrict__ w, int C) {
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {  // sw[tap][c] = bf16(w[c][tap]) (the forward rounds its weights)
    const int tap = i / C, c = i - tap * C;
    sw[i] = bf16_round(w[(size_t)c * 9 + tap]);
  }
}

__global__ void __launch_bounds__(kDw8Threads)
dwconv3x3s2_fwd8_kernel(const bf16* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                        const int32_t* __restrict__ len_out, bf16* __restrict__ out, int T1, int F1, int C, int T2, int F2,
                        int rows_per_cta) {
  extern __shared__ float sw[];  // [9][C]
  load_dw_weights(sw, w, C);
  __syncthreads();
  const int b = blockIdx.y, C8 = C >> 3;
  const int cg = threadIdx.x % C8, pl = threadIdx.x / C8, npl = blockDim.x / C8;
  const int c = cg * 8;
  const int r0 = blockIdx.x * rows_per_cta, nrows = min(rows_per_cta, T2 - r0);
  const int lout = len_out ? len_out[b] : T2;
  float bs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) bs[k] = bias[c + k];
  for (int p = pl; p < nrows * F2; p += npl) {
    const int t2 = r0 + p / F2, f2 = p % F2;
    uint4 xv[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int t = 2 * t2 - 1 + tap / 3, f = 2 * f2 - 1 + tap % 3;
      xv[tap] = (t >= 0 && t < T1 && f >= 0 && f < F1)
                    ? __ldg(reinterpret_cast<const uint4*>(in + (((size_t)b * T1 + t) * F1 + f) * C + c))
                    : make_uint4(0, 0, 0, 0);
    }
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      float x[8];
      unpack8f(xv[tap], x);
      const float4 wa = *reinterpret_cast<const float4*>(sw + tap * C + c), wb = *reinterpret_cast<const float4*>(sw + tap * C + c + 4);
      acc[0] = fmaf(wa.x, x[0], acc[0]); acc[1] = fmaf(wa.y, x[1], acc[1]); acc[2] = fmaf(wa.z, x[2], acc[2]); acc[3] = fmaf(wa.w, x[3], acc[3]);
      acc[4] = fmaf(wb.x, x[4], acc[4]); acc[5] = fmaf(wb.y, x[5], acc[5]); acc[6] = fmaf(wb.z, x[6], acc[6]); acc[7] = fmaf(wb.w, x[7], acc[7]);
    }
    uint4 o = make_uint4(0, 0, 0, 0);
    if (t2 < lout)
      o = make_uint4(pack_bf16(acc[0] + bs[0], acc[1] + bs[1]), pack_bf16(acc[2] + bs[2], acc[3] + bs[3]),
                     pack_bf16(acc[4] + bs[4], acc[5] + bs[5]), pack_bf16(acc[6] + bs[6], acc[7] + bs[7]));
    *reinterpret_cast<uint4*>(out + (((size_t)b * T2 + t2) * F2 + f2) * C + c) = o;
  }
}

__global__ void __launch_bounds__(kDw8Threads)
dwconv3x3s2_bwd_data8_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ w,
                             const int32_t* __restrict__ len_out, bf16* __restrict__ dx, int T1, int F1, int C, int T2, int F2,
                             int rows_per_cta) {
  extern __shared__ float sw[];  // [9][C]
  load_dw_weights(sw, w, C);
  __syncthreads();
  const int b = blockIdx.y, C8 = C >> 3;
  const int cg = threadIdx.x % C8, pl = threadIdx.x / C8, npl = blockDim.x / C8;
  const int c = cg * 8;
  const int r0 = blockIdx.x * rows_per_cta, nrows = min(rows_per_cta, T1 - r0);
  const int lout = len_out ? min(len_out[b], T2) : T2;
  for (int p = pl; p < nrows * F1; p += npl) {
    const int t1 = r0 + p / F1, f1 = p % F1;
    const size_t o = (((size_t)b * T1 + t1) * F1 + f1) * C + c;
    const uint4 xg = __ldg(reinterpret_cast<const uint4*>(x + o));
    // at most 2 x 2 taps reach this input position: kh with (t1 + 1 - kh) even, kw with (f1 + 1 - kw) even
    uint4 gv[4];
    int tapi[4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int kh = ((t1 + 1) & 1) + 2 * a, kw = ((f1 + 1) & 1) + 2 * e;  // kh, kw in {0, 2} or {1}
        const int tn = t1 + 1 - kh, fn = f1 + 1 - kw;
        const bool ok = kh < 3 && kw < 3 && tn >= 0 && fn >= 0 && (tn >> 1) < lout && (fn >> 1) < F2;
        tapi[a * 2 + e] = ok ? kh * 3 + kw : -1;
        gv[a * 2 + e] = ok ? __ldg(reinterpret_cast<const uint4*>(dy + (((size_t)b * T2 + (tn >> 1)) * F2 + (fn >> 1)) * C + c))
                           : make_uint4(0, 0, 0, 0);
      }
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (tapi[q] < 0) continue;
      float g[8];
      unpack8f(gv[q], g);
      const float4 wa = *reinterpret_cast<const float4*>(sw + tapi[q] * C + c), wb = *reinterpret_cast<const float4*>(sw + tapi[q] * C + c + 4);
      acc[0] = fmaf(wa.x, g[0], acc[0]); acc[1] = fmaf(wa.y, g[1], acc[1]); acc[2] = fmaf(wa.z, g[2], acc[2]); acc[3] = fmaf(wa.w, g[3], acc[3]);
      acc[4] = fmaf(wb.x, g[4], acc[4]); acc[5] = fmaf(wb.y, g[5], acc[5]); acc[6] = fmaf(wb.z, g[6], acc[6]); acc[7] = fmaf(wb.w, g[7], acc[7]);
    }
    float xf[8];
    unpack8f(xg, xf);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = xf[k] > 0.f ? acc[k] : 0.f;
    *reinterpret_cast<uint4*>(dx + o) = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]),
                                                   pack_bf16(acc[6], acc[7]));
  }
}

// dw[c, tap] += sum dy * x(tap), dbias[c] += sum dy: per-thread accumulators for 8 channels x (9 taps + bias), summed over the
// CTA's position lanes in shared memory, then one global atomic per (channel, tap) per CTA
__global__ void __launch_bounds__(kDw8Threads)
dwconv3x3s2_bwd_weight8_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const int32_t* __restrict__ len_out,
                               float* __restrict__ dw, float* __restrict__ dbias, int T1, int F1, int C, int T2, int F2,
                               int rows_per_cta) {
  extern __shared__ float red[];  // [C][10]
  for (int i = threadIdx.x; i < C * 10; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int b = blockIdx.y, C8 = C >> 3;
  const int cg = threadIdx.x % C8, pl = threadIdx.x / C8, npl = blockDim.x / C8;
  const int c = cg * 8;
  const int lout = len_out ? min(len_out[b], T2) : T2;
  const int r0 = blockIdx.x * rows_per_cta, nrows = min(rows_per_cta, lout - r0);
  float g[10][8];
#pragma unroll
  for (int j = 0; j < 10; ++j)
#pragma unroll
    for (int k = 0; k < 8; ++k) g[j][k] = 0.f;
  for (int p = pl; p < nrows * F2; p += npl) {
    const int t2 = r0 + p / F2, f2 = p % F2;
    const uint4 gv = __ldg(reinterpret_cast<const uint4*>(dy + (((size_t)b * T2 + t2) * F2 + f2) * C + c));
    uint4 xv[9];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int t = 2 * t2 - 1 + tap / 3, f = 2 * f2 - 1 + tap % 3;
      xv[tap] = (t >= 0 && t < T1 && f >= 0 && f < F1)
                    ? __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * T1 + t) * F1 + f) * C + c))
                    : make_uint4(0, 0, 0, 0);
    }
    float gy[8];
    unpack8f(gv, gy);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[9][k] += gy[k];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      float xf[8];
      unpack8f(xv[tap], xf);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[tap][k] = fmaf(gy[k], xf[k], g[tap][k]);
    }
  }
  if (nrows > 0) {
#pragma unroll
    for (int j = 0; j < 10; ++j)
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(&red[(c + k) * 10 + j], g[j][k]);
  }
  __syncthreads();
  if (nrows > 0)
    for (int i = threadIdx.x; i < C * 10; i += blockDim.x) {
      const int ch = i / 10, j = i - ch * 10;
      if (j < 9) atomicAdd(dw + (size_t)ch * 9 + j, red[i]);
      else atomicAdd(dbias + ch, red[i]);
    }
}

static bool dw_use_vec8(int C) {
  static int forced = -1;  // B2C_DW_IMPL=scalar keeps the 2-channel kernels (A/B comparison)
  if (forced < 0) {
    const char* e = getenv("B2C_DW_IMPL");
    forced = (e && !strcmp(e, "scalar")) ? 1 : 0;
  }
  return forced == 0 && C % 8 == 0 && (C / 8) <= kDw8Threads && kDw8Threads % (C / 8) == 0 && C <= 1024;
}

}  // namespace b2c


using namespace b2c;

extern "C" int b2c_subsample_conv1_fwd(const float* mel, const float* w, const float* bias, const int32_t* len_in,
                                       const int32_t* len_out, void* out_bf16, int B, int F, int T, int C,
                                       void* stream) {
  B2C_REQUIRE(C % 2 == 0 && B > 0 && T > 0 && F > 0, "subsample_conv1_fwd: bad shape");
  const int T1 = (T + 2 - 3) / 2 + 1, F1 = (F + 2 - 3) / 2 + 1;
  subsample_conv1_fwd_kernel<<<dim3(T1, B), 256, 3 * (F + 2) * sizeof(float), (cudaStream_t)stream>>>(
      mel, w, bias, len_in, len_out, reinterpret_cast<bf16*>(out_bf16), F, T, C, T1, F1);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_subsample_im2col(const float* mel, const int32_t* len_in, void* patches_bf16, int B, int F, int T,
                                    void* stream) {
  B2C_REQUIRE(B > 0 && T > 0 && F > 0, "subsample_im2col: bad shape");
  const int T1 = (T + 2 - 3) / 2 + 1, F1 = (F + 2 - 3) / 2 + 1;
  const size_t smem = (size_t)F * (2 * kImT + 3) * sizeof(float);
  B2C_REQUIRE(smem <= 48 * 1024, "subsample_im2col: F = %d too large", F);
  subsample_im2col_kernel<<<dim3((T1 + kImT - 1) / kImT, B), 256, smem, (cudaStream_t)stream>>>(
      mel, len_in, reinterpret_cast<bf16*>(patches_bf16), F, T, T1, F1);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_subsample_conv1_bwd(const float* mel, const void* dout_bf16, const int32_t* len_in, float* dw,
                                       float* dbias, int B, int F, int T, int C, void* stream) {
  B2C_REQUIRE(C % 2 == 0 && B > 0 && T > 0 && F > 0, "subsample_conv1_bwd: bad shape");
  const int T1 = (T + 2 - 3) / 2 + 1, F1 = (F + 2 - 3) / 2 + 1;
  const int rows = 16;
  subsample_conv1_bwd_kernel<<<dim3((T1 + rows - 1) / rows, B), 256, 3 * (F + 2) * sizeof(float), (cudaStream_t)stream>>>(
      mel, reinterpret_cast<const bf16*>(dout_bf16), len_in, dw, dbias, F, T, C, T1, F1, rows);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_dwconv3x3s2_fwd(const void* in_bf16, const float* w, const float* bias, const int32_t* len_out,
                                   void* out_bf16, int B, int T1, int F1, int C, void* stream) {
  B2C_REQUIRE(C % 2 == 0 && B > 0 && T1 > 0 && F1 > 0, "dwconv3x3s2_fwd: bad shape");
  const int T2 = (T1 + 2 - 3) / 2 + 1, F2 = (F1 + 2 - 3) / 2 + 1;
  if (dw_use_vec8(C)) {
    const int rows = 4;
    dwconv3x3s2_fwd8_kernel<<<dim3((T2 + rows - 1) / rows, B), kDw8Threads, 9 * C * sizeof(float), (cudaStream_t)stream>>>(
        reinterpret_cast<const bf16*>(in_bf16), w, bias, len_out, reinterpret_cast<bf16*>(out_bf16), T1, F1, C, T2, F2, rows);
    B2C_LAUNCH_CHECK();
    return B2C_OK;
  }
  dwconv3x3s2_fwd_kernel<<<dim3(T2, B), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(in_bf16), w, bias, len_out, reinterpret_cast<bf16*>(out_bf16), T1, F1, C, T2, F2);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_dwconv3x3s2_bwd(const void* dy_bf16, const void* x_bf16, const float* w, const int32_t* len_out,
                                   void* dx_bf16, float* dw, float* dbias, int B, int T1, int F1, int C, void* stream) {
  B2C_REQUIRE(C % 2 == 0 && B > 0 && T1 > 0 && F1 > 0, "dwconv3x3s2_bwd: bad shape");
  const int T2 = (T1 + 2 - 3) / 2 + 1, F2 = (F1 + 2 - 3) / 2 + 1;
  if (dw_use_vec8(C)) {
    const int rows_d = 4, rows_w = 16;
    dwconv3x3s2_bwd_data8_kernel<<<dim3((T1 + rows_d - 1) / rows_d, B), kDw8Threads, 9 * C * sizeof(float), (cudaStream_t)stream>>>(
        reinterpret_cast<const bf16*>(dy_bf16), reinterpret_cast<const bf16*>(x_bf16), w, len_out, reinterpret_cast<bf16*>(dx_bf16), T1,
        F1, C, T2, F2, rows_d);
    B2C_LAUNCH_CHECK();
    dwconv3x3s2_bwd_weight8_kernel<<<dim3((T2 + rows_w - 1) / rows_w, B), kDw8Threads, 10 * C * sizeof(float), (cudaStream_t)stream>>>(
        reinterpret_cast<const bf16*>(dy_bf16), reinterpret_cast<const bf16*>(x_bf16), len_out, dw, dbias, T1, F1, C, T2, F2, rows_w);
    B2C_LAUNCH_CHECK();
    return B2C_OK;
  }
  dwconv3x3s2_bwd_data_kernel<<<dim3(T1, B), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(dy_bf16), reinterpret_cast<const bf16*>(x_bf16), w, len_out,
      reinterpret_cast<bf16*>(dx_bf16), T1, F1, C, T2, F2);
  B2C_LAUNCH_CHECK();
  const int rows = 8;
  dwconv3x3s2_bwd_weight_kernel<<<dim3((T2 + rows - 1) / rows, B), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(dy_bf16), reinterpret_cast<const bf16*>(x_bf16), len_out, dw, dbias, T1, F1, C, T2, F2,
      rows);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
