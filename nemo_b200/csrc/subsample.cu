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
// (input is already time-masked by its producer).  thread = (channel pair, row slice); CTA = kDwRows time rows of one batch
// item, so the 18 strided weight loads of a thread are amortised over kDwRows / rs rows x F positions (one row per CTA made
// that prologue a third of the kernel; splitting the f axis over more threads instead made it worse — measured).
// With C = 256 (FastConformer) the 128 channel pairs leave room for rs = 2 row slices in a 256-thread CTA.
constexpr int kDwThreads = 256;
constexpr int kDwRows = 8;

struct DwMap {
  int CP, rs, rpart, cp0, cstep;
  bool idle;
};
__device__ __forceinline__ DwMap dw_map(int C) {
  DwMap m;
  m.CP = C >> 1;
  if (m.CP >= (int)blockDim.x) {  // more channel pairs than threads: loop over them, one row slice
    m.rs = 1; m.rpart = 0; m.cp0 = (int)threadIdx.x; m.cstep = (int)blockDim.x; m.idle = false;
  } else {
    m.rs = (int)blockDim.x / m.CP; m.rpart = (int)threadIdx.x / m.CP; m.cp0 = (int)threadIdx.x % m.CP; m.cstep = m.CP;
    m.idle = m.rpart >= m.rs;
  }
  return m;
}

__global__ void __launch_bounds__(kDwThreads)
dwconv3x3s2_fwd_kernel(const bf16* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                       const int32_t* __restrict__ len_out, bf16* __restrict__ out, int T1, int F1, int C, int T2, int F2) {
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * kDwRows, r1 = min(T2, r0 + kDwRows);
  const int lout = len_out ? len_out[b] : T2;
  const DwMap m = dw_map(C);
  if (m.idle) return;
  for (int cp = m.cp0; cp < m.CP; cp += m.cstep) {
    const int c = cp * 2;
    float w0[9], w1[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      w0[j] = bf16_round(w[(size_t)c * 9 + j]);
      w1[j] = bf16_round(w[(size_t)(c + 1) * 9 + j]);
    }
    const float b0 = bias[c], b1 = bias[c + 1];
    for (int t2 = r0 + m.rpart; t2 < r1; t2 += m.rs) {
      const bool valid = t2 < lout;
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
}

// depthwise 3x3 s2 backward, data: dx[b,t1,f1,c] = (x[b,t1,f1,c] > 0) * sum_{kh,kw} dy[b,(t1+1-kh)/2,(f1+1-kw)/2,c] * w[c,kh,kw]
// (dy rows t2 >= len_out count as zero: the forward masks the depthwise output before the pointwise conv)
__global__ void __launch_bounds__(kDwThreads)
dwconv3x3s2_bwd_data_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const float* __restrict__ w,
                            const int32_t* __restrict__ len_out, bf16* __restrict__ dx, int T1, int F1, int C, int T2,
                            int F2) {
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * kDwRows, r1 = min(T1, r0 + kDwRows);
  const int lout = len_out ? len_out[b] : T2;
  const DwMap m = dw_map(C);
  if (m.idle) return;
  for (int cp = m.cp0; cp < m.CP; cp += m.cstep) {
    const int c = cp * 2;
    float w0[9], w1[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      w0[j] = bf16_round(w[(size_t)c * 9 + j]);
      w1[j] = bf16_round(w[(size_t)(c + 1) * 9 + j]);
    }
    for (int t1 = r0 + m.rpart; t1 < r1; t1 += m.rs) {
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
}

// weight / bias gradients of the depthwise 3x3 s2 conv: dw[c,kh,kw] += sum dy[b,t2,f2,c] * x[b,2t2+kh-1,2f2+kw-1,c]
// (CP * rs == blockDim.x is arranged by the launcher; the row slices of a channel pair are summed through shared memory, then
// one atomic per (channel, tap) per CTA)
__global__ void __launch_bounds__(kDwThreads)
dwconv3x3s2_bwd_weight_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x, const int32_t* __restrict__ len_out,
                              float* __restrict__ dw, float* __restrict__ dbias, int T1, int F1, int C, int T2, int F2,
                              int rows_per_cta) {
  __shared__ float red[kDwThreads * 20];
  const int b = blockIdx.y;
  const int lout = len_out ? min(len_out[b], T2) : T2;
  const int t2_0 = blockIdx.x * rows_per_cta, t2_1 = min(lout, t2_0 + rows_per_cta);
  const int CP = C >> 1, rs = (int)blockDim.x / CP;
  const int rpart = (int)threadIdx.x / CP, cp = (int)threadIdx.x % CP;
  const int c = cp * 2;
  float g0[9], g1[9], d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int j = 0; j < 9; ++j) g0[j] = g1[j] = 0.f;
  for (int t2 = t2_0 + rpart; t2 < t2_1; t2 += rs) {
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
  float* mine = red + threadIdx.x * 20;
#pragma unroll
  for (int j = 0; j < 9; ++j) { mine[j] = g0[j]; mine[9 + j] = g1[j]; }
  mine[18] = d0; mine[19] = d1;
  __syncthreads();
  if (rpart == 0 && t2_1 > t2_0) {
    float acc[20];
#pragma unroll
    for (int j = 0; j < 20; ++j) acc[j] = 0.f;
    for (int q = 0; q < rs; ++q)
#pragma unroll
      for (int j = 0; j < 20; ++j) acc[j] += red[(q * CP + cp) * 20 + j];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      atomicAdd(dw + (size_t)c * 9 + j, acc[j]);
      atomicAdd(dw + (size_t)(c + 1) * 9 + j, acc[9 + j]);
    }
    atomicAdd(dbias + c, acc[18]);
    atomicAdd(dbias + c + 1, acc[19]);
  }
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
  dwconv3x3s2_fwd_kernel<<<dim3((T2 + kDwRows - 1) / kDwRows, B), kDwThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(in_bf16), w, bias, len_out, reinterpret_cast<bf16*>(out_bf16), T1, F1, C, T2, F2);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_dwconv3x3s2_bwd(const void* dy_bf16, const void* x_bf16, const float* w, const int32_t* len_out,
                                   void* dx_bf16, float* dw, float* dbias, int B, int T1, int F1, int C, void* stream) {
  B2C_REQUIRE(C % 2 == 0 && B > 0 && T1 > 0 && F1 > 0, "dwconv3x3s2_bwd: bad shape");
  const int T2 = (T1 + 2 - 3) / 2 + 1, F2 = (F1 + 2 - 3) / 2 + 1;
  dwconv3x3s2_bwd_data_kernel<<<dim3((T1 + kDwRows - 1) / kDwRows, B), kDwThreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(dy_bf16), reinterpret_cast<const bf16*>(x_bf16), w, len_out,
      reinterpret_cast<bf16*>(dx_bf16), T1, F1, C, T2, F2);
  B2C_LAUNCH_CHECK();
  // CTA size = the largest multiple of C/2 (channel pairs) that fits kDwThreads; 32 output rows per CTA
  B2C_REQUIRE(C / 2 <= kDwThreads, "dwconv3x3s2_bwd: C = %d too large", C);
  const int rows = 32;
  const int wthreads = (kDwThreads / (C / 2)) * (C / 2);
  dwconv3x3s2_bwd_weight_kernel<<<dim3((T2 + rows - 1) / rows, B), wthreads, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(dy_bf16), reinterpret_cast<const bf16*>(x_bf16), len_out, dw, dbias, T1, F1, C, T2, F2,
      rows);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
