// HBM-bound row/column kernels of the Conformer block: LayerNorm fwd/bwd (+ residual-gradient fan-in, dropout of the
// outgoing branch gradient), bias-gradient column sums, fp32→bf16 weight casts / permutes, rel-pos table, lengths.
// One warp per row, 128-bit loads, warp-shuffle reductions.  Reference: torch.nn.LayerNorm(d) x5 per ConformerLayer
// (conformer_modules.py:94,98,108,154,158) under bf16 autocast (fp32 maths, bf16 consumer).
#include "common.cuh"

namespace b2c {

B2C_DEFINE_SEED_SETTER(set_seed_counter_elementwise)


constexpr int kMaxVecLimit = 8;  // d <= 1024 (8 float4 per lane)

template <int kMaxVec>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     bf16* __restrict__ y16, float* __restrict__ y32, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int M, int d) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int nv = d >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)warp * d);
  float4 v[kMaxVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = i * 32 + lane;
    v[i] = (c < nv) ? xr[c] : make_float4(0, 0, 0, 0);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = i * 32 + lane;
    if (c < nv) {
      float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, dd = v[i].w - mean;
      q += a * a + b * b + cc * cc + dd * dd;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
  if (lane == 0) {
    if (mean_out) mean_out[warp] = mean;
    if (rstd_out) rstd_out[warp] = rstd;
  }
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = i * 32 + lane;
    if (c < nv) {
      const float4 g = reinterpret_cast<const float4*>(gamma)[c];
      const float4 bt = reinterpret_cast<const float4*>(beta)[c];
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + bt.x;
      o.y = (v[i].y - mean) * rstd * g.y + bt.y;
      o.z = (v[i].z - mean) * rstd * g.z + bt.z;
      o.w = (v[i].w - mean) * rstd * g.w + bt.w;
      if (y32) reinterpret_cast<float4*>(y32 + (size_t)warp * d)[c] = o;
      if (y16) reinterpret_cast<uint2*>(y16 + (size_t)warp * d)[c] = make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
    }
  }
}

struct DropArgs {
  unsigned long long seed;
  unsigned int site, keep;
  float scale;
};

// rows handled per warp are strided so each warp keeps per-column dgamma/dbeta partials in registers
template <bool DY_BF16, int kMaxVec>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const void* __restrict__ dy_, const float* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dr,
                     int accumulate, float* __restrict__ dgamma, float* __restrict__ dbeta, bf16* __restrict__ next16,
                     float next_scale, DropArgs drop, int M, int d) {
  extern __shared__ float red[];  // [2][d] CTA partials
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int nv = d >> 2;
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float4 ag[kMaxVec], ab[kMaxVec];
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) ag[i] = ab[i] = make_float4(0, 0, 0, 0);
  auto load_dy = [&](int row, int c) -> float4 {
    if (DY_BF16) {
      const uint2 u = reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dy_) + (size_t)row * d)[c];
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
      return make_float4(a.x, a.y, b.x, b.y);
    }
    return reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy_) + (size_t)row * d)[c];
  };
  for (int row = blockIdx.x * wpb + wib; row < M; row += gridDim.x * wpb) {
    const float mu = mean[row], rs = rstd[row];
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * d);
    // pass 1: row sums for dx and the dgamma / dbeta partials (nothing is kept in registers across the reduction;
    // pass 2 re-reads x / dy from L1 — fewer live registers → more resident warps → more bytes in flight)
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = i * 32 + lane;
      if (c < nv) {
        const float4 xv = xr[c];
        const float4 dv = load_dy(row, c);
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const float4 xh = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        ag[i].x += dv.x * xh.x; ag[i].y += dv.y * xh.y; ag[i].z += dv.z * xh.z; ag[i].w += dv.w * xh.w;
        ab[i].x += dv.x; ab[i].y += dv.y; ab[i].z += dv.z; ab[i].w += dv.w;
        const float a0 = dv.x * gm.x, a1 = dv.y * gm.y, a2 = dv.z * gm.z, a3 = dv.w * gm.w;
        s1 += a0 + a1 + a2 + a3;
        s2 += a0 * xh.x + a1 * xh.y + a2 * xh.z + a3 * xh.w;
      }
    }
    s1 = warp_sum(s1) / (float)d;
    s2 = warp_sum(s2) / (float)d;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      const int c = i * 32 + lane;
      if (c < nv) {
        const float4 xv = xr[c];
        const float4 dv = load_dy(row, c);
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        float4 o;
        o.x = rs * (dv.x * gm.x - s1 - (xv.x - mu) * rs * s2);
        o.y = rs * (dv.y * gm.y - s1 - (xv.y - mu) * rs * s2);
        o.z = rs * (dv.z * gm.z - s1 - (xv.z - mu) * rs * s2);
        o.w = rs * (dv.w * gm.w - s1 - (xv.w - mu) * rs * s2);
        float4* drp = reinterpret_cast<float4*>(dr + (size_t)row * d) + c;
        if (accumulate) {
          const float4 p = *drp;
          o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        *drp = o;
        if (next16) {
          float n0 = o.x * next_scale, n1 = o.y * next_scale, n2 = o.z * next_scale, n3 = o.w * next_scale;
          if (drop.keep != 0) {
            const uint64_t idx = (uint64_t)row * (uint64_t)d + (uint64_t)c * 4;
            const uint32_t hb = dropout_base(drop.seed, drop.site, idx >> 4);
            const uint32_t khi = dropout_keep_hi(drop.keep);
            uint32_t w0, w1, w2, w3;
            dropout_pair(hb, ((uint32_t)idx >> 1) & 7u, w0, w1);
            dropout_pair(hb, (((uint32_t)idx >> 1) & 7u) + 1u, w2, w3);
            n0 = (w0 < khi) ? n0 * drop.scale : 0.f;
            n1 = (w1 < khi) ? n1 * drop.scale : 0.f;
            n2 = (w2 < khi) ? n2 * drop.scale : 0.f;
            n3 = (w3 < khi) ? n3 * drop.scale : 0.f;
          }
          reinterpret_cast<uint2*>(next16 + (size_t)row * d)[c] = make_uint2(pack_bf16(n0, n1), pack_bf16(n2, n3));
        }
      }
    }
  }
  // CTA reduction of dgamma/dbeta partials, then one atomic per column per CTA
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    const int c = i * 32 + lane;
    if (c < nv) {
      atomicAdd(&red[c * 4 + 0], ag[i].x); atomicAdd(&red[c * 4 + 1], ag[i].y);
      atomicAdd(&red[c * 4 + 2], ag[i].z); atomicAdd(&red[c * 4 + 3], ag[i].w);
      atomicAdd(&red[d + c * 4 + 0], ab[i].x); atomicAdd(&red[d + c * 4 + 1], ab[i].y);
      atomicAdd(&red[d + c * 4 + 2], ab[i].z); atomicAdd(&red[d + c * 4 + 3], ab[i].w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, red[i]);
    if (dbeta) atomicAdd(dbeta + i, red[d + i]);
  }
}

// out[n] += sum_m in[m, n]; CTA = 256 columns (32 lanes x 8 bf16 = 16 B per lane) x row-chunk, 4 rows in flight per warp
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const bf16* __restrict__ in, float* __restrict__ out, int M, int N, int ld, int rows_per_cta) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  const bool vec = (n + 8 <= N) && (ld % 8 == 0);
  if (vec) {
    int m = r0 + w;
    for (; m + 24 < r1; m += 32) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(in + (size_t)(m + 8 * k) * ld + n);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 p0 = unpack_bf16(u[k].x), p1 = unpack_bf16(u[k].y), p2 = unpack_bf16(u[k].z), p3 = unpack_bf16(u[k].w);
        a[0] += p0.x; a[1] += p0.y; a[2] += p1.x; a[3] += p1.y; a[4] += p2.x; a[5] += p2.y; a[6] += p3.x; a[7] += p3.y;
      }
    }
    for (; m < r1; m += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(in + (size_t)m * ld + n);
      float2 p0 = unpack_bf16(u.x), p1 = unpack_bf16(u.y), p2 = unpack_bf16(u.z), p3 = unpack_bf16(u.w);
      a[0] += p0.x; a[1] += p0.y; a[2] += p1.x; a[3] += p1.y; a[4] += p2.x; a[5] += p2.y; a[6] += p3.x; a[7] += p3.y;
    }
  } else {
    for (int m = r0 + w; m < r1; m += 8)
      for (int j = 0; j < 8; ++j)
        if (n + j < N) a[j] += __bfloat162float(in[(size_t)m * ld + n + j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[w][lane * 8 + j] = a[j];
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i][threadIdx.x];
  const int nn = blockIdx.x * 256 + threadIdx.x;
  if (nn < N) atomicAdd(out + nn, sum);
}

__global__ void cast_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n, float scale) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 v = *reinterpret_cast<const float4*>(in + i);
    *reinterpret_cast<uint2*>(out + i) = make_uint2(pack_bf16(v.x * scale, v.y * scale), pack_bf16(v.z * scale, v.w * scale));
  } else {
    for (; i < n; ++i) out[i] = __float2bfloat16_rn(in[i] * scale);
  }
}

// out_bf16 = bf16(scale * dropout(in)) — gradient entering pre_encode.out (xscale + pre-encoder dropout backward)
__global__ void scale_drop_cast_kernel(const float* __restrict__ in, bf16* __restrict__ out, long long n, float scale,
                                       DropArgs drop) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = in[i] * scale;
  if (drop.keep != 0) v = dropout_keep(drop.seed, drop.site, (uint64_t)i, drop.keep) ? v * drop.scale : 0.f;
  out[i] = __float2bfloat16_rn(v);
}

// SpecAugment (spectr_augment.py:135-215): out[b,f,t] = mask_value where t falls in one of the utterance's n_t time spans
// or f in one of its n_f frequency bands, else in[b,f,t].  One pass; the reference builds [B,n,axis] mask tensors and runs
// two masked_fill passes.  Span bounds come from the host module (same random draws as the reference).
__global__ void __launch_bounds__(256)
spec_augment_kernel(const float* __restrict__ in, float* __restrict__ out, const int32_t* __restrict__ t_start,
                    const int32_t* __restrict__ t_end, const int32_t* __restrict__ f_start, const int32_t* __restrict__ f_end,
                    int F, int T, int n_t, int n_f, float mask_value) {
  extern __shared__ int32_t sb[];  // [2*n_t] time spans, [2*n_f] frequency bands of this utterance
  const int b = blockIdx.z, f = blockIdx.y;
  for (int i = threadIdx.x; i < n_t; i += blockDim.x) { sb[2 * i] = t_start[b * n_t + i]; sb[2 * i + 1] = t_end[b * n_t + i]; }
  for (int i = threadIdx.x; i < n_f; i += blockDim.x) {
    sb[2 * n_t + 2 * i] = f_start[b * n_f + i];
    sb[2 * n_t + 2 * i + 1] = f_end[b * n_f + i];
  }
  __syncthreads();
  bool fmask = false;
  for (int i = 0; i < n_f; ++i) fmask |= (f >= sb[2 * n_t + 2 * i] && f < sb[2 * n_t + 2 * i + 1]);
  const size_t row = ((size_t)b * F + f) * T;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    bool m = fmask;
    for (int i = 0; i < n_t; ++i) m |= (t >= sb[2 * i] && t < sb[2 * i + 1]);
    out[row + t] = m ? mask_value : in[row + t];
  }
}

// in [A,B,C] → out; mode 0: out[b][a][c] (swap dims 0,1); mode 1: out[a][c][b] (swap dims 1,2)
template <typename TO, bool ACCUM>
__global__ void permute3_kernel(const float* __restrict__ in, TO* __restrict__ out, int A, int B, int C, int mode) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)A * B * C;
  if (i >= n) return;
  // i indexes the OUTPUT (coalesced writes)
  long long src;
  if (mode == 0) {
    const int c = (int)(i % C);
    const int a = (int)((i / C) % A);
    const int b = (int)(i / ((long long)C * A));
    src = ((long long)a * B + b) * C + c;
  } else {
    const int b = (int)(i % B);
    const int c = (int)((i / B) % C);
    const int a = (int)(i / ((long long)B * C));
    src = ((long long)a * B + b) * C + c;
  }
  if constexpr (sizeof(TO) == 2) {
    out[i] = __float2bfloat16_rn(in[src]);
  } else {
    if (ACCUM) out[i] += in[src];
    else out[i] = in[src];
  }
}

__global__ void fill_f32_kernel(float* p, long long n, float v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// rows r = 0..2T-2 ↔ position T-1-r; pe[2i] = sin(pos*div_i), pe[2i+1] = cos(pos*div_i)
// (multi_head_attention.py:1007-1020: div = exp(-2i * ln(10000)/d) computed in fp32)
__global__ void relpos_table_kernel(bf16* __restrict__ pe16, float* __restrict__ pe32, int T, int d) {
  const int r = blockIdx.x;
  const float pos = (float)(T - 1 - r);
  for (int i = threadIdx.x; i < d / 2; i += blockDim.x) {
    const float div = expf((float)(2 * i) * -(9.210340371976184f / (float)d));
    const float a = pos * div;
    const float s = sinf(a), c = cosf(a);
    if (pe16) {
      pe16[(size_t)r * d + 2 * i] = __float2bfloat16_rn(s);
      pe16[(size_t)r * d + 2 * i + 1] = __float2bfloat16_rn(c);
    }
    if (pe32) {
      pe32[(size_t)r * d + 2 * i] = s;
      pe32[(size_t)r * d + 2 * i + 1] = c;
    }
  }
}


// ---- CTC head (ConvASRDecoder, nemo/collections/asr/modules/conv_asr.py:456-468): row log-softmax of bf16 logits
// one warp per row; logits [M, ld] bf16 (first V columns valid) → log-probs [M, V] fp32
__global__ void __launch_bounds__(256)
log_softmax_fwd_kernel(const bf16* __restrict__ logits, int ld, float* __restrict__ out, int M, int V) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const bf16* x = logits + (size_t)row * ld;
  float mx = -3.0e38f;
  for (int j = lane; j < V; j += 32) mx = fmaxf(mx, __bfloat162float(x[j]));
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < V; j += 32) s += __expf(__bfloat162float(x[j]) - mx);
  const float lse = mx + __logf(warp_sum(s));
  float* o = out + (size_t)row * V;
  for (int j = lane; j < V; j += 32) o[j] = __bfloat162float(x[j]) - lse;
}
// dlogits[m, j] = dlogp[m, j] - exp(logp[m, j]) * sum_j dlogp[m, j]   (bf16, pad columns [V, ld) zeroed)
__global__ void __launch_bounds__(256)
log_softmax_bwd_kernel(const float* __restrict__ dlogp, const float* __restrict__ logp, bf16* __restrict__ dlogits, int ld,
                       int M, int V) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* g = dlogp + (size_t)row * V;
  const float* lp = logp + (size_t)row * V;
  float s = 0.f;
  for (int j = lane; j < V; j += 32) s += g[j];
  s = warp_sum(s);
  bf16* o = dlogits + (size_t)row * ld;
  for (int j = lane; j < ld; j += 32) o[j] = __float2bfloat16_rn(j < V ? g[j] - __expf(lp[j]) * s : 0.f);
}

// calc_length / calculate_conv_output_size with k=3, s=2, pad 1+1: L → (L + 2 - 3)//2 + 1 (floor), n times
__global__ void subsample_lengths_kernel(const int64_t* __restrict__ len_in, int32_t* __restrict__ len_in32,
                                         int32_t* __restrict__ len_stage, int64_t* __restrict__ len_out64, int B,
                                         int n_convs) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  long long l = len_in[b];
  if (len_in32) len_in32[b] = (int32_t)l;
  for (int i = 0; i < n_convs; ++i) {
    long long num = l + 2 - 3;
    long long q = num / 2;
    if (num % 2 != 0 && num < 0) --q;  // floor
    l = q + 1;
    if (len_stage) len_stage[i * B + b] = (int32_t)l;
  }
  if (len_out64) len_out64[b] = l;
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32,
                                 float* mean, float* rstd, int M, int d, void* stream) {
  B2C_REQUIRE(d % 4 == 0 && d <= kMaxVecLimit * 128, "layernorm: d=%d unsupported (need d%%4==0, d<=1024)", d);
  B2C_REQUIRE(M > 0, "layernorm: M=%d", M);
  const int wpb = 8;
  const dim3 grid((M + wpb - 1) / wpb);
  bf16* y16 = reinterpret_cast<bf16*>(y_bf16);
  cudaStream_t st = (cudaStream_t)stream;
  if (d <= 256) layernorm_fwd_kernel<2><<<grid, wpb * 32, 0, st>>>(x, gamma, beta, y16, y_f32, mean, rstd, M, d);
  else if (d <= 512) layernorm_fwd_kernel<4><<<grid, wpb * 32, 0, st>>>(x, gamma, beta, y16, y_f32, mean, rstd, M, d);
  else layernorm_fwd_kernel<8><<<grid, wpb * 32, 0, st>>>(x, gamma, beta, y16, y_f32, mean, rstd, M, d);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_layernorm_bwd(const void* dy, int dy_is_bf16, const float* x, const float* gamma, const float* mean,
                                 const float* rstd, float* dr, int accumulate, float* dgamma, float* dbeta,
                                 void* next_bf16, float next_scale, float drop_p, unsigned long long seed,
                                 unsigned int site, int M, int d, void* stream) {
  B2C_REQUIRE(d % 4 == 0 && d <= kMaxVecLimit * 128, "layernorm_bwd: d=%d unsupported", d);
  DropArgs da{seed, site, 0u, 1.f};
  if (drop_p > 0.f) {
    da.keep = (unsigned int)((1.0 - (double)drop_p) * 65536.0 + 0.5);
    da.scale = 1.f / (1.f - drop_p);
  }
  const int wpb = 8;
  int grid = (M + wpb - 1) / wpb;
  const int cap = num_sms() * (d <= 512 ? 3 : 2);  // = resident CTAs: one wave, rows strided across it
  if (grid > cap) grid = cap;
  const size_t smem = 2 * (size_t)d * sizeof(float);
  bf16* nx = reinterpret_cast<bf16*>(next_bf16);
  cudaStream_t st = (cudaStream_t)stream;
#define B2C_LN_BWD(BF, NV)                                                                                          \
  layernorm_bwd_kernel<BF, NV><<<grid, wpb * 32, smem, st>>>(dy, x, gamma, mean, rstd, dr, accumulate, dgamma, dbeta, nx, \
                                                            next_scale, da, M, d)
  if (dy_is_bf16) {
    if (d <= 256) B2C_LN_BWD(true, 2); else if (d <= 512) B2C_LN_BWD(true, 4); else B2C_LN_BWD(true, 8);
  } else {
    if (d <= 256) B2C_LN_BWD(false, 2); else if (d <= 512) B2C_LN_BWD(false, 4); else B2C_LN_BWD(false, 8);
  }
#undef B2C_LN_BWD
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_colsum_bf16(const void* in_bf16, float* out, int M, int N, int ld, void* stream) {
  B2C_REQUIRE(M > 0 && N > 0 && ld % 2 == 0, "colsum: bad args");
  const int col_blocks = (N + 255) / 256;
  int row_blocks = (num_sms() * 4 + col_blocks - 1) / col_blocks;
  if (row_blocks < 1) row_blocks = 1;
  int rows_per_cta = (M + row_blocks - 1) / row_blocks;
  if (rows_per_cta < 64) rows_per_cta = 64;
  row_blocks = (M + rows_per_cta - 1) / rows_per_cta;
  colsum_bf16_kernel<<<dim3(col_blocks, row_blocks), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const bf16*>(in_bf16), out, M, N, ld, rows_per_cta);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_cast_bf16(const float* in, void* out_bf16, long long n, float scale, void* stream) {
  B2C_REQUIRE(n > 0, "cast: n=%lld", n);
  const long long threads = (n + 3) / 4;
  cast_bf16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      in, reinterpret_cast<bf16*>(out_bf16), n, scale);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_scale_drop_cast(const float* in, void* out_bf16, long long n, float scale, float drop_p,
                                   unsigned long long seed, unsigned int site, void* stream) {
  B2C_REQUIRE(n > 0, "scale_drop_cast: n=%lld", n);
  DropArgs da{seed, site, 0u, 1.f};
  if (drop_p > 0.f) {
    da.keep = (unsigned int)((1.0 - (double)drop_p) * 65536.0 + 0.5);
    da.scale = 1.f / (1.f - drop_p);
  }
  scale_drop_cast_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      in, reinterpret_cast<bf16*>(out_bf16), n, scale, da);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_permute3_bf16(const float* in, void* out_bf16, int A, int B, int C, int mode, void* stream) {
  const long long n = (long long)A * B * C;
  B2C_REQUIRE(n > 0 && (mode == 0 || mode == 1), "permute3: bad args");
  permute3_kernel<bf16, false><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      in, reinterpret_cast<bf16*>(out_bf16), A, B, C, mode);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_permute3_f32(const float* in, float* out, int A, int B, int C, int mode, int accumulate,
                                void* stream) {
  const long long n = (long long)A * B * C;
  B2C_REQUIRE(n > 0 && (mode == 0 || mode == 1), "permute3: bad args");
  if (accumulate)
    permute3_kernel<float, true><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, A, B, C, mode);
  else
    permute3_kernel<float, false><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, A, B, C, mode);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_fill_f32(float* p, long long n, float v, void* stream) {
  if (n <= 0) return B2C_OK;
  fill_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, n, v);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_relpos_table(void* pos_bf16, float* pos_f32, int T, int d, void* stream) {
  B2C_REQUIRE(T > 0 && d % 2 == 0, "relpos_table: bad args");
  relpos_table_kernel<<<2 * T - 1, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<bf16*>(pos_bf16), pos_f32, T, d);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_subsample_lengths(const int64_t* len_in, int32_t* len_in32, int32_t* len_stage, int64_t* len_out64,
                                     int B, int n_convs, void* stream) {
  B2C_REQUIRE(B > 0 && n_convs >= 0, "subsample_lengths: bad args");
  subsample_lengths_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(len_in, len_in32, len_stage, len_out64, B,
                                                                             n_convs);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_log_softmax_fwd(const void* logits_bf16, int ld, float* out, int M, int V, void* stream) {
  B2C_REQUIRE(M > 0 && V > 0 && ld >= V, "log_softmax_fwd: bad args");
  log_softmax_fwd_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const bf16*>(logits_bf16), ld, out, M, V);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_log_softmax_bwd(const float* dlogp, const float* logp, void* dlogits_bf16, int ld, int M, int V,
                                   void* stream) {
  B2C_REQUIRE(M > 0 && V > 0 && ld >= V, "log_softmax_bwd: bad args");
  log_softmax_bwd_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(dlogp, logp, reinterpret_cast<bf16*>(dlogits_bf16),
                                                                         ld, M, V);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_spec_augment(const float* in, float* out, const int32_t* t_start, const int32_t* t_end,
                                const int32_t* f_start, const int32_t* f_end, int B, int F, int T, int n_t, int n_f,
                                float mask_value, void* stream) {
  B2C_REQUIRE(B > 0 && F > 0 && T > 0 && n_t >= 0 && n_f >= 0 && n_t + n_f <= 4096, "spec_augment: bad shape");
  const int tb = (T + 255) / 256;
  spec_augment_kernel<<<dim3(tb < 8 ? tb : 8, F, B), 256, 2 * (n_t + n_f) * sizeof(int32_t), (cudaStream_t)stream>>>(
      in, out, t_start, t_end, f_start, f_end, F, T, n_t, n_f, mask_value);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
