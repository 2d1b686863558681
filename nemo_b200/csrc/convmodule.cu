// K6: ConformerConvolution core between the two pointwise GEMMs (conformer_modules.py:320-350):
//   depthwise k-tap conv over time (CausalConv1D with symmetric padding, causal_convs.py:130-150) + bias,
//   BatchNorm1d (train: batch statistics over all B*T frames incl. padded ones; eval: running stats), SiLU,
//   and their backward passes.  Channels-last [B*T, d] bf16 tiles staged in shared memory with a (k-1) halo,
//   two channels per thread (bf16x2), per-thread register accumulators, one atomic per channel per CTA.
#include "common.cuh"

namespace b2c {

constexpr int kTT = 32;         // output time steps per tile
constexpr int kCh = 256;        // channels per CTA (128 threads x 2)

// ------------------------------------------------------------------------------------------- forward
template <int K>
__global__ void __launch_bounds__(128)
dwconv_fwd_kernel(const bf16* __restrict__ u, const float* __restrict__ w, const float* __restrict__ bias,
                  bf16* __restrict__ c, float* __restrict__ bn_sum, float* __restrict__ bn_sumsq, int T, int d,
                  int tiles_per_cta) {
  extern __shared__ uint32_t su[];  // [(kTT + K - 1)][128] bf16x2
  constexpr int PAD = (K - 1) / 2;
  const int b = blockIdx.z;
  const int ch = blockIdx.x * kCh + threadIdx.x * 2;
  const bool act = ch < d;
  float w0[K], w1[K];
  float b0 = 0.f, b1 = 0.f;
  if (act) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      w0[j] = w[(size_t)ch * K + j];
      w1[j] = w[(size_t)(ch + 1) * K + j];
    }
    b0 = bias ? bias[ch] : 0.f;
    b1 = bias ? bias[ch + 1] : 0.f;
  }
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  const bf16* ub = u + (size_t)b * T * d;
  bf16* cb = c + (size_t)b * T * d;
  for (int tile = 0; tile < tiles_per_cta; ++tile) {
    const int t0 = (blockIdx.y * tiles_per_cta + tile) * kTT;
    if (t0 >= T) break;
    __syncthreads();
#pragma unroll 16
    for (int r = 0; r < kTT + K - 1; ++r) {  // unrolled so the independent global loads are in flight together
      const int t = t0 + r - PAD;
      uint32_t v = 0u;
      if (act && t >= 0 && t < T) v = __ldg(reinterpret_cast<const uint32_t*>(ub + (size_t)t * d + ch));
      su[r * 128 + threadIdx.x] = v;
    }
    __syncthreads();
    if (act) {
      const int nt = min(kTT, T - t0);
      // register blocking: 4 consecutive outputs share every staged input row (each row is read / unpacked once)
      for (int i = 0; i < nt; i += 4) {
        float a0[4], a1[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) { a0[o] = b0; a1[o] = b1; }
#pragma unroll
        for (int r = 0; r < K + 3; ++r) {
          const float2 x = unpack_bf16(su[(i + r) * 128 + threadIdx.x]);
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const int j = r - o;
            if (j >= 0 && j < K) {
              a0[o] = fmaf(w0[j], x.x, a0[o]);
              a1[o] = fmaf(w1[j], x.y, a1[o]);
            }
          }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          if (i + o < nt) {
            const uint32_t packed = pack_bf16(a0[o], a1[o]);
            *reinterpret_cast<uint32_t*>(cb + (size_t)(t0 + i + o) * d + ch) = packed;
            const float2 r2 = unpack_bf16(packed);  // BatchNorm sees the bf16 conv output
            s0 += r2.x; s1 += r2.y; q0 += r2.x * r2.x; q1 += r2.y * r2.y;
          }
        }
      }
    }
  }
  if (act && bn_sum) {
    atomicAdd(bn_sum + ch, s0); atomicAdd(bn_sum + ch + 1, s1);
    atomicAdd(bn_sumsq + ch, q0); atomicAdd(bn_sumsq + ch + 1, q1);
  }
}

__global__ void bn_finalize_kernel(const float* __restrict__ bn_sum, const float* __restrict__ bn_sumsq,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out, float count, int d, int training, float momentum,
                                   float eps, long long* __restrict__ num_batches_tracked) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && num_batches_tracked) *num_batches_tracked += 1;  // nn.BatchNorm1d bookkeeping
  if (c >= d) return;
  float mean, var;
  if (training) {
    mean = bn_sum[c] / count;
    var = fmaxf(bn_sumsq[c] / count - mean * mean, 0.f);  // biased variance normalises (nn.BatchNorm1d)
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  if (mean_out) mean_out[c] = mean;
  if (rstd_out) rstd_out[c] = rstd;
}

// s = bf16(silu(bf16(c*scale + shift)))   [M, d], d even
__global__ void __launch_bounds__(256)
bn_silu_fwd_kernel(const bf16* __restrict__ c, const float* __restrict__ scale, const float* __restrict__ shift,
                   bf16* __restrict__ s, long long total_pairs, int d) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_pairs) return;
  const int ch = (int)((i * 2) % d);
  const float2 x = unpack_bf16(reinterpret_cast<const uint32_t*>(c)[i]);
  const float n0 = bf16_round(fmaf(x.x, scale[ch], shift[ch]));
  const float n1 = bf16_round(fmaf(x.y, scale[ch + 1], shift[ch + 1]));
  reinterpret_cast<uint32_t*>(s)[i] = pack_bf16(silu_f(n0), silu_f(n1));
}

// ------------------------------------------------------------------------------------------- backward
// phase 1: per-channel sum(dz), sum(dz*xhat) with dz = ds * silu'(bn), bn = bf16(c*scale+shift)
__global__ void __launch_bounds__(128)
bn_silu_bwd_reduce_kernel(const bf16* __restrict__ ds, const bf16* __restrict__ c, const float* __restrict__ scale,
                          const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ rstd,
                          float* __restrict__ sum_dz, float* __restrict__ sum_dz_xhat, long long M, int d,
                          int rows_per_cta) {
  const int ch = blockIdx.x * kCh + threadIdx.x * 2;
  if (ch >= d) return;
  const float sc0 = scale[ch], sc1 = scale[ch + 1], sh0 = shift[ch], sh1 = shift[ch + 1];
  const float m0 = mean[ch], m1 = mean[ch + 1], r0 = rstd[ch], r1 = rstd[ch + 1];
  const long long row0 = (long long)blockIdx.y * rows_per_cta;
  const long long row1 = min(M, row0 + rows_per_cta);
  float a0 = 0.f, a1 = 0.f, x0 = 0.f, x1 = 0.f;
  for (long long r = row0; r < row1; r += 4) {  // 8 independent loads in flight per thread
    uint32_t cw[4], dw_[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool ok = r + k < row1;
      cw[k] = ok ? __ldg(reinterpret_cast<const uint32_t*>(c + (r + k) * d + ch)) : 0u;
      dw_[k] = ok ? __ldg(reinterpret_cast<const uint32_t*>(ds + (r + k) * d + ch)) : 0u;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 cv = unpack_bf16(cw[k]), dv = unpack_bf16(dw_[k]);  // padded rows: ds = 0 → contribute nothing
      const float n0 = bf16_round(fmaf(cv.x, sc0, sh0)), n1 = bf16_round(fmaf(cv.y, sc1, sh1));
      const float dz0 = dv.x * dsilu_f(n0), dz1 = dv.y * dsilu_f(n1);
      a0 += dz0; a1 += dz1;
      x0 += dz0 * (cv.x - m0) * r0;
      x1 += dz1 * (cv.y - m1) * r1;
    }
  }
  atomicAdd(sum_dz + ch, a0); atomicAdd(sum_dz + ch + 1, a1);
  atomicAdd(sum_dz_xhat + ch, x0); atomicAdd(sum_dz_xhat + ch + 1, x1);
}

// phase 2: dc = gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat)) (train) | gamma*rstd*dz (eval)
__global__ void __launch_bounds__(256)
bn_silu_bwd_apply_kernel(const bf16* __restrict__ ds, const bf16* __restrict__ c, const float* __restrict__ gamma,
                         const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                         const float* __restrict__ rstd, const float* __restrict__ sum_dz,
                         const float* __restrict__ sum_dz_xhat, bf16* __restrict__ dc, float* __restrict__ dgamma,
                         float* __restrict__ dbeta, long long total_pairs, int d, float inv_count, int training) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < d && dgamma) {  // parameter gradients once (first d threads of the grid)
    atomicAdd(dgamma + i, sum_dz_xhat[i]);
    atomicAdd(dbeta + i, sum_dz[i]);
  }
  if (i >= total_pairs) return;
  const int ch = (int)((i * 2) % d);
  const float2 cv = unpack_bf16(reinterpret_cast<const uint32_t*>(c)[i]);
  const float2 dv = unpack_bf16(reinterpret_cast<const uint32_t*>(ds)[i]);
  float o[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int cc = ch + k;
    const float cval = k ? cv.y : cv.x, dval = k ? dv.y : dv.x;
    const float n = bf16_round(fmaf(cval, scale[cc], shift[cc]));
    const float dz = dval * dsilu_f(n);
    const float gr = gamma[cc] * rstd[cc];
    if (training) {
      const float xhat = (cval - mean[cc]) * rstd[cc];
      o[k] = gr * (dz - sum_dz[cc] * inv_count - xhat * sum_dz_xhat[cc] * inv_count);
    } else {
      o[k] = gr * dz;
    }
  }
  reinterpret_cast<uint32_t*>(dc)[i] = pack_bf16(o[0], o[1]);
}

// du[t] = mask(t) * sum_j w[j] * dc[t + PAD - j];  dw[j] += sum_t dc[t] * u[t + j - PAD];  dbias += sum_t dc[t]
template <int K>
__global__ void __launch_bounds__(128)
dwconv_bwd_kernel(const bf16* __restrict__ dc, const bf16* __restrict__ u, const float* __restrict__ w,
                  const int32_t* __restrict__ lengths, bf16* __restrict__ du, float* __restrict__ dw,
                  float* __restrict__ dbias, int T, int d, int tiles_per_cta) {
  extern __shared__ uint32_t sm[];
  constexpr int PAD = (K - 1) / 2;
  constexpr int ROWS = kTT + K - 1;
  uint32_t* sdc = sm;               // [ROWS][128]
  uint32_t* suu = sm + ROWS * 128;  // [ROWS][128]
  const int b = blockIdx.z;
  const int ch = blockIdx.x * kCh + threadIdx.x * 2;
  const bool act = ch < d;
  const int len = lengths ? lengths[b] : T;
  float w0[K], w1[K], g0[K], g1[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    w0[j] = act ? w[(size_t)ch * K + j] : 0.f;
    w1[j] = act ? w[(size_t)(ch + 1) * K + j] : 0.f;
    g0[j] = g1[j] = 0.f;
  }
  float db0 = 0.f, db1 = 0.f;
  const bf16* dcb = dc + (size_t)b * T * d;
  const bf16* ub = u + (size_t)b * T * d;
  bf16* dub = du + (size_t)b * T * d;
  for (int tile = 0; tile < tiles_per_cta; ++tile) {
    const int t0 = (blockIdx.y * tiles_per_cta + tile) * kTT;
    if (t0 >= T) break;
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < ROWS; ++r) {
      const int t = t0 + r - PAD;
      uint32_t a = 0u, bb = 0u;
      if (act && t >= 0 && t < T) {
        a = __ldg(reinterpret_cast<const uint32_t*>(dcb + (size_t)t * d + ch));
        bb = __ldg(reinterpret_cast<const uint32_t*>(ub + (size_t)t * d + ch));
      }
      sdc[r * 128 + threadIdx.x] = a;
      suu[r * 128 + threadIdx.x] = bb;
    }
    __syncthreads();
    if (act) {
      const int nt = min(kTT, T - t0);
      for (int i = 0; i < nt; i += 4) {
        // ---- du[t0+i+o] = sum_j w[j] * dc[(t0+i+o) + PAD - j]  ↔ smem row (i + o + 2*PAD - j); 4 outputs per row pass
        float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < K + 3; ++r) {  // r = smem row - i
          const float2 x = unpack_bf16(sdc[(i + r) * 128 + threadIdx.x]);
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const int j = o + 2 * PAD - r;
            if (j >= 0 && j < K) {
              a0[o] = fmaf(w0[j], x.x, a0[o]);
              a1[o] = fmaf(w1[j], x.y, a1[o]);
            }
          }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o)
          if (i + o < nt) {
            const bool valid = (t0 + i + o) < len;
            *reinterpret_cast<uint32_t*>(dub + (size_t)(t0 + i + o) * d + ch) = valid ? pack_bf16(a0[o], a1[o]) : 0u;
          }
        // ---- dw[j] += dc[t0+i+o] (smem row i+o+PAD) * u[t0+i+o+j-PAD] (smem row i+o+j); db += dc
        float2 g[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          g[o] = (i + o < nt) ? unpack_bf16(sdc[(i + o + PAD) * 128 + threadIdx.x]) : make_float2(0.f, 0.f);
          db0 += g[o].x; db1 += g[o].y;
        }
#pragma unroll
        for (int r = 0; r < K + 3; ++r) {  // u smem row i + r feeds tap j = r - o of output o
          const float2 x = unpack_bf16(suu[(i + r) * 128 + threadIdx.x]);
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const int j = r - o;
            if (j >= 0 && j < K) {
              g0[j] = fmaf(g[o].x, x.x, g0[j]);
              g1[j] = fmaf(g[o].y, x.y, g1[j]);
            }
          }
        }
      }
    }
  }
  if (act) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      atomicAdd(dw + (size_t)ch * K + j, g0[j]);
      atomicAdd(dw + (size_t)(ch + 1) * K + j, g1[j]);
    }
    if (dbias) { atomicAdd(dbias + ch, db0); atomicAdd(dbias + ch + 1, db1); }
  }
}

// interleaved (value, gate): dg[2j] = du*sig(g); dg[2j+1] = du*a*sig(g)*(1-sig(g))
__global__ void __launch_bounds__(256)
glu_bwd_kernel(const bf16* __restrict__ du, const bf16* __restrict__ g, bf16* __restrict__ dg, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // index into du [M*d]
  if (i >= total) return;
  const float dv = __bfloat162float(du[i]);
  const float2 ag = unpack_bf16(reinterpret_cast<const uint32_t*>(g)[i]);
  const float s = sigmoid_f(ag.y);
  reinterpret_cast<uint32_t*>(dg)[i] = pack_bf16(dv * s, dv * ag.x * s * (1.f - s));
}

// ---- 8-wide variants (d % 8 == 0): one 16-byte load / store per tensor and thread — the 4-byte-per-thread versions above
// keep too few bytes in flight to get near the HBM roofline
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__global__ void __launch_bounds__(256)
bn_silu_fwd8_kernel(const bf16* __restrict__ c, const float* __restrict__ scale, const float* __restrict__ shift,
                    bf16* __restrict__ s, long long total8, int d) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int ch = (int)((i * 8) % d);
  float x[8], sc[8], sh[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(c) + i), x);
  ld8(scale + ch, sc);
  ld8(shift + ch, sh);
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = silu_f(bf16_round(fmaf(x[k], sc[k], sh[k])));
  reinterpret_cast<uint4*>(s)[i] = pack8(x);
}

__global__ void __launch_bounds__(256)
bn_silu_bwd_apply8_kernel(const bf16* __restrict__ ds, const bf16* __restrict__ c, const float* __restrict__ gamma,
                          const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                          const float* __restrict__ rstd, const float* __restrict__ sum_dz,
                          const float* __restrict__ sum_dz_xhat, bf16* __restrict__ dc, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, long long total8, int d, float inv_count, int training) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < d && dgamma) {  // parameter gradients once (first d threads of the grid; the host checks total8 >= d)
    atomicAdd(dgamma + i, sum_dz_xhat[i]);
    atomicAdd(dbeta + i, sum_dz[i]);
  }
  if (i >= total8) return;
  const int ch = (int)((i * 8) % d);
  float cv[8], dv[8], sc[8], sh[8], gm[8], rs[8], mu[8], s1[8], s2[8], o[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(c) + i), cv);
  unpack8(__ldg(reinterpret_cast<const uint4*>(ds) + i), dv);
  ld8(scale + ch, sc); ld8(shift + ch, sh); ld8(gamma + ch, gm); ld8(rstd + ch, rs);
  if (training) { ld8(mean + ch, mu); ld8(sum_dz + ch, s1); ld8(sum_dz_xhat + ch, s2); }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float n = bf16_round(fmaf(cv[k], sc[k], sh[k]));
    const float dz = dv[k] * dsilu_f(n);
    const float gr = gm[k] * rs[k];
    if (training) {
      const float xhat = (cv[k] - mu[k]) * rs[k];
      o[k] = gr * (dz - s1[k] * inv_count - xhat * s2[k] * inv_count);
    } else {
      o[k] = gr * dz;
    }
  }
  reinterpret_cast<uint4*>(dc)[i] = pack8(o);
}

// 8 du elements ↔ 16 interleaved (value, gate) elements of g / dg
__global__ void __launch_bounds__(256)
glu_bwd8_kernel(const bf16* __restrict__ du, const bf16* __restrict__ g, bf16* __restrict__ dg, long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  float dv[8], g0[8], g1[8], o0[8], o1[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(du) + i), dv);
  unpack8(__ldg(reinterpret_cast<const uint4*>(g) + 2 * i), g0);
  unpack8(__ldg(reinterpret_cast<const uint4*>(g) + 2 * i + 1), g1);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float s0 = sigmoid_f(g0[2 * k + 1]), s1 = sigmoid_f(g1[2 * k + 1]);
    o0[2 * k] = dv[k] * s0;
    o0[2 * k + 1] = dv[k] * g0[2 * k] * s0 * (1.f - s0);
    o1[2 * k] = dv[4 + k] * s1;
    o1[2 * k + 1] = dv[4 + k] * g1[2 * k] * s1 * (1.f - s1);
  }
  reinterpret_cast<uint4*>(dg)[2 * i] = pack8(o0);
  reinterpret_cast<uint4*>(dg)[2 * i + 1] = pack8(o1);
}

// Time tiles per CTA so that the grid runs in as few "tile-times" as possible: waves(grid / resident slots) * tiles per CTA
// (a grid of 1.15 waves costs two), most tiles per CTA (fewest weight reloads / gradient atomics) on ties.
static int pick_tiles_per_cta(long long ctas_per_tile_row, int tiles, long long slots) {
  int best = 1;
  long long best_t = -1;
  for (int tpc = 1; tpc <= tiles; ++tpc) {
    const long long grid = ctas_per_tile_row * ((tiles + tpc - 1) / tpc);
    const long long t = ((grid + slots - 1) / slots) * tpc;
    if (best_t < 0 || t <= best_t) { best_t = t; best = tpc; }
  }
  return best;
}

template <int K>
static int launch_dw_fwd(const bf16* u, const float* w, const float* bias, bf16* c, float* s, float* q, int B, int T,
                         int d, cudaStream_t st) {
  const int tiles = (T + kTT - 1) / kTT;
  const int cb = (d + kCh - 1) / kCh;
  const size_t smem = (size_t)(kTT + K - 1) * 128 * 4;
  static int occ = 0;
  if (!occ) B2C_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dwconv_fwd_kernel<K>, 128, smem));
  const int tpc = pick_tiles_per_cta(cb * B, tiles, num_sms() * (occ > 0 ? occ : 1));
  dwconv_fwd_kernel<K><<<dim3(cb, (tiles + tpc - 1) / tpc, B), 128, smem, st>>>(u, w, bias, c, s, q, T, d, tpc);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

template <int K>
static int launch_dw_bwd(const bf16* dc, const bf16* u, const float* w, const int32_t* len, bf16* du, float* dw,
                         float* db, int B, int T, int d, cudaStream_t st) {
  const int tiles = (T + kTT - 1) / kTT;
  const int cb = (d + kCh - 1) / kCh;
  const size_t smem = (size_t)(kTT + K - 1) * 128 * 4 * 2;
  static int occ = 0;
  if (!occ) {
    B2C_CUDA(cudaFuncSetAttribute(dwconv_bwd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2C_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dwconv_bwd_kernel<K>, 128, smem));
  }
  const int tpc = pick_tiles_per_cta(cb * B, tiles, num_sms() * (occ > 0 ? occ : 1));
  dwconv_bwd_kernel<K><<<dim3(cb, (tiles + tpc - 1) / tpc, B), 128, smem, st>>>(dc, u, w, len, du, dw, db, T, d, tpc);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_dwconv_fwd(const void* u_bf16, const float* w, const float* bias, void* c_bf16, float* bn_sum,
                              float* bn_sumsq, int B, int T, int d, int k, void* stream) {
  B2C_REQUIRE(d % 2 == 0 && B > 0 && T > 0, "dwconv_fwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const bf16* u = reinterpret_cast<const bf16*>(u_bf16);
  bf16* c = reinterpret_cast<bf16*>(c_bf16);
  switch (k) {
    case 31: return launch_dw_fwd<31>(u, w, bias, c, bn_sum, bn_sumsq, B, T, d, st);
    case 9: return launch_dw_fwd<9>(u, w, bias, c, bn_sum, bn_sumsq, B, T, d, st);
    case 5: return launch_dw_fwd<5>(u, w, bias, c, bn_sum, bn_sumsq, B, T, d, st);
  }
  set_last_error("dwconv_fwd: kernel size %d not built (31, 9, 5)", k);
  return B2C_ERR_UNSUPPORTED;
}

extern "C" int b2c_bn_finalize(const float* bn_sum, const float* bn_sumsq, const float* gamma, const float* beta,
                               float* running_mean, float* running_var, float* scale, float* shift, float* mean_out,
                               float* rstd_out, long long count, int d, int training, float momentum, float eps,
                               long long* num_batches_tracked, void* stream) {
  bn_finalize_kernel<<<(d + 127) / 128, 128, 0, (cudaStream_t)stream>>>(bn_sum, bn_sumsq, gamma, beta, running_mean,
                                                                       running_var, scale, shift, mean_out, rstd_out,
                                                                       (float)count, d, training, momentum, eps,
                                                                       num_batches_tracked);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_bn_silu_fwd(const void* c_bf16, const float* scale, const float* shift, void* s_bf16, long long M,
                               int d, void* stream) {
  B2C_REQUIRE(d % 2 == 0, "bn_silu: d odd");
  const long long pairs = M * d / 2;
  if (d % 8 == 0 && (reinterpret_cast<uintptr_t>(c_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(s_bf16) & 15) == 0) {
    const long long n8 = M * d / 8;
    bn_silu_fwd8_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const bf16*>(c_bf16), scale, shift, reinterpret_cast<bf16*>(s_bf16), n8, d);
    B2C_LAUNCH_CHECK();
    return B2C_OK;
  }
  bn_silu_fwd_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(c_bf16), scale, shift, reinterpret_cast<bf16*>(s_bf16), pairs, d);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_bn_silu_bwd_reduce(const void* ds_bf16, const void* c_bf16, const float* scale, const float* shift,
                                      const float* mean, const float* rstd, float* sum_dz, float* sum_dz_xhat,
                                      long long M, int d, void* stream) {
  const int cb = (d + kCh - 1) / kCh;
  int row_blocks = (num_sms() * 8 + cb - 1) / cb;
  int rows_per_cta = (int)((M + row_blocks - 1) / row_blocks);
  if (rows_per_cta < 32) rows_per_cta = 32;
  row_blocks = (int)((M + rows_per_cta - 1) / rows_per_cta);
  bn_silu_bwd_reduce_kernel<<<dim3(cb, row_blocks), 128, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(ds_bf16), reinterpret_cast<const bf16*>(c_bf16), scale, shift, mean, rstd, sum_dz,
      sum_dz_xhat, M, d, rows_per_cta);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_bn_silu_bwd_apply(const void* ds_bf16, const void* c_bf16, const float* gamma, const float* scale,
                                     const float* shift, const float* mean, const float* rstd, const float* sum_dz,
                                     const float* sum_dz_xhat, void* dc_bf16, float* dgamma, float* dbeta, long long M,
                                     int d, int training, void* stream) {
  const long long pairs = M * d / 2;
  B2C_REQUIRE(pairs >= d, "bn_silu_bwd_apply: tensor smaller than d");
  if (d % 8 == 0 && M >= 8 && (reinterpret_cast<uintptr_t>(c_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(ds_bf16) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dc_bf16) & 15) == 0) {
    const long long n8 = M * d / 8;  // >= d because M >= 8
    bn_silu_bwd_apply8_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const bf16*>(ds_bf16), reinterpret_cast<const bf16*>(c_bf16), gamma, scale, shift, mean, rstd,
        sum_dz, sum_dz_xhat, reinterpret_cast<bf16*>(dc_bf16), dgamma, dbeta, n8, d, 1.0f / (float)M, training);
    B2C_LAUNCH_CHECK();
    return B2C_OK;
  }
  bn_silu_bwd_apply_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(ds_bf16), reinterpret_cast<const bf16*>(c_bf16), gamma, scale, shift, mean, rstd,
      sum_dz, sum_dz_xhat, reinterpret_cast<bf16*>(dc_bf16), dgamma, dbeta, pairs, d, 1.0f / (float)M, training);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_dwconv_bwd(const void* dc_bf16, const void* u_bf16, const float* w, const int32_t* lengths,
                              void* du_bf16, float* dw, float* dbias, int B, int T, int d, int k, void* stream) {
  B2C_REQUIRE(d % 2 == 0 && B > 0 && T > 0, "dwconv_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const bf16* dc = reinterpret_cast<const bf16*>(dc_bf16);
  const bf16* u = reinterpret_cast<const bf16*>(u_bf16);
  bf16* du = reinterpret_cast<bf16*>(du_bf16);
  switch (k) {
    case 31: return launch_dw_bwd<31>(dc, u, w, lengths, du, dw, dbias, B, T, d, st);
    case 9: return launch_dw_bwd<9>(dc, u, w, lengths, du, dw, dbias, B, T, d, st);
    case 5: return launch_dw_bwd<5>(dc, u, w, lengths, du, dw, dbias, B, T, d, st);
  }
  set_last_error("dwconv_bwd: kernel size %d not built (31, 9, 5)", k);
  return B2C_ERR_UNSUPPORTED;
}

extern "C" int b2c_glu_bwd(const void* du_bf16, const void* g_bf16, void* dg_bf16, long long M, int d, void* stream) {
  const long long total = M * d;
  if (total % 8 == 0 && (reinterpret_cast<uintptr_t>(du_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(g_bf16) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dg_bf16) & 15) == 0) {
    glu_bwd8_kernel<<<(unsigned)((total / 8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const bf16*>(du_bf16), reinterpret_cast<const bf16*>(g_bf16), reinterpret_cast<bf16*>(dg_bf16),
        total / 8);
    B2C_LAUNCH_CHECK();
    return B2C_OK;
  }
  glu_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(du_bf16), reinterpret_cast<const bf16*>(g_bf16), reinterpret_cast<bf16*>(dg_bf16),
      total);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
