// K1: fused mel front-end — FilterbankFeatures.forward (nemo/collections/asr/parts/preprocessing/features.py:423-502).
//
//   kernel 1 (mel_frames):   dither → pre-emphasis → length mask → centred framing (hop 160, n_fft 512) →
//                            hann(400) window → 512-pt FFT in shared memory (two real frames packed into one
//                            complex radix-8x8x8 transform, in place in one 4 KB buffer per warp) → |X|^2 → sparse Slaney
//                            mel → log(x + 2^-24) → [B, n_mels, T] via a smem-staged, time-coalesced store.
//                            16 frames per CTA, 56 KB of shared memory: 4 CTAs (32 warps) per SM.
//   kernel 2 (mel_normalize): per (b, mel) exact two-pass mean / unbiased std over frames < seq_len
//                            (normalize_batch, features.py:59-93), normalise, zero frames >= seq_len.
//
// HBM roofline: algorithmic bytes = 4*S in + 4*n_mels*T out per utterance (96 kB per audio-second).
// The intermediate [B,257,T] complex spectrum of the reference (197 MB at 32x30 s) never exists.
#include "common.cuh"

namespace b2c {

B2C_DEFINE_SEED_SETTER(set_seed_counter_mel)


constexpr int kNfft = 512;
constexpr int kFramesPerCta = 16;
constexpr int kMelWarps = 8;
constexpr int kPad = 512 + 16;  // 512 + 512/32 padding floats

__device__ __forceinline__ int padi(int i) { return i + (i >> 5); }

struct cpx {
  float re, im;
};
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return {a.re - b.re, a.im - b.im}; }
// multiply by -i  (forward-transform rotation by -90 deg)
__device__ __forceinline__ cpx mul_mi(cpx a) { return {a.im, -a.re}; }

// in-register 8-point forward DFT, natural order in and out
__device__ __forceinline__ void dft8(cpx (&v)[8]) {
  const float h = 0.70710678118654752440f;
  cpx a0 = cadd(v[0], v[4]), a1 = csub(v[0], v[4]);
  cpx a2 = cadd(v[2], v[6]), a3 = mul_mi(csub(v[2], v[6]));
  cpx a4 = cadd(v[1], v[5]), a5 = csub(v[1], v[5]);
  cpx a6 = cadd(v[3], v[7]), a7 = mul_mi(csub(v[3], v[7]));
  cpx b0 = cadd(a0, a2), b2 = csub(a0, a2);
  cpx b1 = cadd(a1, a3), b3 = csub(a1, a3);
  cpx b4 = cadd(a4, a6), b6 = mul_mi(csub(a4, a6));
  cpx b5 = cadd(a5, a7), b7 = csub(a5, a7);
  // twiddles W8^1 = (1-i)/sqrt2 on b5, W8^3 = (-1-i)/sqrt2 on b7
  cpx t5 = {h * (b5.re + b5.im), h * (b5.im - b5.re)};
  cpx t7 = {h * (b7.im - b7.re), -h * (b7.re + b7.im)};
  v[0] = cadd(b0, b4); v[4] = csub(b0, b4);
  v[2] = cadd(b2, b6); v[6] = csub(b2, b6);
  v[1] = cadd(b1, t5); v[5] = csub(b1, t5);
  v[3] = cadd(b3, t7); v[7] = csub(b3, t7);
}

__device__ __forceinline__ float gauss_from_philox(uint32_t a, uint32_t b) {
  // Box-Muller on two 32-bit uniforms
  float u1 = ((float)a + 1.0f) * (1.0f / 4294967296.0f);
  float u2 = (float)b * (1.0f / 4294967296.0f);
  return sqrtf(-2.f * __logf(u1)) * __cosf(6.283185307179586f * u2);
}

__global__ void mel_seq_len_kernel(const int64_t* __restrict__ length, int64_t* __restrict__ seq_len, int B, int n_fft,
                                   int hop) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  // features.py:413-417: floor((len + 2*(n_fft//2) - n_fft) / hop); len == 0 → 0 (:424-427)
  int64_t len = length[b];
  int64_t num = len + (int64_t)(n_fft / 2) * 2 - n_fft;
  int64_t q = num / hop;
  if ((num % hop != 0) && ((num < 0) != (hop < 0))) --q;  // floor division
  seq_len[b] = (len == 0) ? 0 : q;
}

// fb_ranges[m] = {first non-zero bin, one-past-last non-zero bin} of filter m
__global__ void mel_fb_ranges_kernel(const float* __restrict__ fb, int n_mels, int n_bins, int2* __restrict__ ranges) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_mels) return;
  int lo = n_bins, hi = 0;
  for (int k = 0; k < n_bins; ++k)
    if (fb[(size_t)m * n_bins + k] != 0.f) {
      lo = min(lo, k);
      hi = max(hi, k + 1);
    }
  if (lo >= hi) { lo = 0; hi = 0; }
  ranges[m] = make_int2(lo, hi);
}

// Occupancy is what bounds this kernel (every warp runs a serial chain of shared-memory FFT stages): the per-warp FFT
// buffer is ONE padded re/im pair used in place (stage inputs are staged through registers), the power spectra reuse it, and
// a CTA covers only kFramesPerCta = 16 frames, so a CTA needs ~57 KB and four of them (32 warps) share an SM.
__global__ void __launch_bounds__(kMelWarps * 32, 4)
mel_frames_kernel(const float* __restrict__ x, const int64_t* __restrict__ length, const float* __restrict__ window,
                  const float* __restrict__ fb, const int2* __restrict__ fb_ranges, float* __restrict__ mel, int S, int T,
                  int n_mels, int win, int hop, float preemph, float dither, unsigned long long seed) {
  extern __shared__ float sm[];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFramesPerCta;
  const int nframes = min(kFramesPerCta, T - t0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_bins = kNfft / 2 + 1;
  const int woff = (kNfft - win) / 2;  // window is centred in the n_fft frame (torch.stft)

  // smem carve-up
  const int ylen = (kFramesPerCta - 1) * hop + win;  // samples covered by the windowed part of the tile's frames
  float* y = sm;                                  // [ylen]
  float* tw_re = y + ((ylen + 3) & ~3);           // [512]
  float* tw_im = tw_re + kNfft;                   // [512]
  float* wsm = tw_im + kNfft;                     // [win]
  float* stage = wsm + ((win + 3) & ~3);          // [n_mels][kFramesPerCta+1]
  float* fftbuf = stage + n_mels * (kFramesPerCta + 1);  // per warp: 2 * kPad (re, im), used in place

  // twiddles e^{-2 pi i k / 512} and window
  for (int k = threadIdx.x; k < kNfft; k += blockDim.x) {
    float s, c;
    sincospif(-(float)k / 256.f, &s, &c);
    tw_re[k] = c;
    tw_im[k] = s;
  }
  for (int k = threadIdx.x; k < win; k += blockDim.x) wsm[k] = window[k];

  // y[n] for n in [n0, n0+ylen): pre-emphasised, masked, (dithered) signal; n0 = t0*hop - n_fft/2 + woff
  const int64_t len = length[b];
  if (dither > 0.f) seed = seed_eff(seed);
  const int64_t n0 = (int64_t)t0 * hop - kNfft / 2 + woff;
  const float* xb = x + (size_t)b * S;
  for (int i = threadIdx.x; i < ylen; i += blockDim.x) {
    const int64_t n = n0 + i;
    float v = 0.f;
    if (n >= 0 && n < S && n < len) {
      float cur = xb[n];
      float prev = (n > 0) ? xb[n - 1] : 0.f;
      if (dither > 0.f) {
        // features.py:435-436 (training only): x += dither * randn.  Stateless Philox keyed by (seed, b, n).
        uint4 r = philox4((uint32_t)n, (uint32_t)b, 0x6d656c31u, 0, (uint32_t)seed, (uint32_t)(seed >> 32));
        cur += dither * gauss_from_philox(r.x, r.y);
        if (n > 0) {
          uint4 r2 = philox4((uint32_t)(n - 1), (uint32_t)b, 0x6d656c31u, 0, (uint32_t)seed, (uint32_t)(seed >> 32));
          prev += dither * gauss_from_philox(r2.x, r2.y);
        }
      }
      v = (n > 0) ? cur - preemph * prev : cur;  // features.py:441
    }
    y[i] = v;
  }
  __syncthreads();

  float* bre = fftbuf + warp * 2 * kPad;
  float* bim = bre + kPad;

  // each warp: frame pairs (2*pp, 2*pp+1) for pp = warp, warp+8, ...
  for (int pp = warp; pp * 2 < kFramesPerCta; pp += kMelWarps) {
    const int fa = pp * 2, fbf = pp * 2 + 1;
    if (fa < nframes) {
      const float* ya = y + fa * hop;
      const float* yb = y + fbf * hop;  // may run past nframes: harmless (result discarded)
      // ---- stage 1 (Ns = 1): read windowed input at j + r*64, write at 8j + r
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int j = lane + half * 32;
        cpx v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int n = j + r * 64 - woff;
          float a = 0.f, bb = 0.f;
          if (n >= 0 && n < win) {
            const float w = wsm[n];
            a = w * ya[n];
            bb = w * yb[n];
          }
          v[r] = {a, bb};
        }
        dft8(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int o = padi(j * 8 + r);
          bre[o] = v[r].re;
          bim[o] = v[r].im;
        }
      }
      __syncwarp();
      // ---- stage 2 (Ns = 8), in place: every lane pulls the inputs of BOTH its butterflies into registers first
      {
        cpx v[2][8];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int j = lane + half * 32;
          const int k = j & 7;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int i = padi(j + r * 64);
            cpx d = {bre[i], bim[i]};
            const int ti = r * k * 8;  // angle = -2 pi r k / 64 = -2 pi (8 r k) / 512
            cpx w = {tw_re[ti], tw_im[ti]};
            v[half][r] = cmul(d, w);
          }
        }
        __syncwarp();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int j = lane + half * 32;
          dft8(v[half]);
          const int j0 = (j >> 3) * 64 + (j & 7);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int o = padi(j0 + r * 8);
            bre[o] = v[half][r].re;
            bim[o] = v[half][r].im;
          }
        }
      }
      __syncwarp();
      // ---- stage 3 (Ns = 64), in place: a butterfly reads and writes the same eight slots (natural order out)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int j = lane + half * 32;
        cpx v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int i = padi(j + r * 64);
          cpx d = {bre[i], bim[i]};
          const int ti = r * j;  // angle = -2 pi r j / 512
          cpx w = {tw_re[ti], tw_im[ti]};
          v[r] = cmul(d, w);
        }
        dft8(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int o = padi(j + r * 64);
          bre[o] = v[r].re;
          bim[o] = v[r].im;
        }
      }
      __syncwarp();
      // ---- split the packed transform: A = (Z[k] + conj Z[-k])/2, B = (Z[k] - conj Z[-k])/(2i); the power spectra go
      // back into the same buffer (pwA = bre[0..256], pwB = bim[0..256], unpadded), so read everything first
      {
        float pa[9], pb[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) {
          const int k = lane + q * 32;
          pa[q] = pb[q] = 0.f;
          if (k < n_bins) {
            const int km = (kNfft - k) & (kNfft - 1);
            const float zr = bre[padi(k)], zi = bim[padi(k)];
            const float cr = bre[padi(km)], ci = -bim[padi(km)];
            const float ar = 0.5f * (zr + cr), ai = 0.5f * (zi + ci);
            const float br = 0.5f * (zi - ci), bi = -0.5f * (zr - cr);
            // features.py:450-461: magnitude (sqrt) then pow(2.0)
            const float ma = sqrtf(ar * ar + ai * ai), mb = sqrtf(br * br + bi * bi);
            pa[q] = ma * ma;
            pb[q] = mb * mb;
          }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 9; ++q) {
          const int k = lane + q * 32;
          if (k < n_bins) {
            bre[k] = pa[q];
            bim[k] = pb[q];
          }
        }
      }
      __syncwarp();
      // ---- sparse mel projection + log (features.py:469-475)
      for (int m = lane; m < n_mels; m += 32) {
        const int2 rg = fb_ranges[m];
        const float* f = fb + (size_t)m * n_bins;
        float sa = 0.f, sb = 0.f;
        for (int k = rg.x; k < rg.y; ++k) {
          const float w = __ldg(f + k);
          sa = fmaf(w, bre[k], sa);
          sb = fmaf(w, bim[k], sb);
        }
        stage[m * (kFramesPerCta + 1) + fa] = logf(sa + 5.9604644775390625e-08f);
        stage[m * (kFramesPerCta + 1) + fbf] = logf(sb + 5.9604644775390625e-08f);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // coalesced store: rows of `nframes` consecutive floats
  float* outb = mel + (size_t)b * n_mels * T + t0;
  for (int i = threadIdx.x; i < n_mels * kFramesPerCta; i += blockDim.x) {
    const int m = i / kFramesPerCta, f = i - m * kFramesPerCta;
    if (f < nframes) outb[(size_t)m * T + f] = stage[m * (kFramesPerCta + 1) + f];
  }
}

// one CTA per (b, mel) row: exact two-pass statistics over valid frames, normalise, mask.
__global__ void __launch_bounds__(256)
mel_normalize_kernel(float* __restrict__ mel, const int64_t* __restrict__ seq_len, int T, int normalize, float pad_value) {
  __shared__ float red[8];
  __shared__ float bcast;
  // grid = (n_mels, B)
  const int64_t n64 = seq_len[blockIdx.y];
  const int n = (int)min((int64_t)T, n64);
  float* row = mel + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mean = 0.f, inv_std = 1.f;
  if (normalize) {
    float s = 0.f;
    for (int t = tid; t < n; t += blockDim.x) s += row[t];
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < 8; ++w) tot += red[w];
      bcast = tot / (float)n;  // n == 0 → NaN like the reference (0/0); masked to pad_value below
    }
    __syncthreads();
    mean = bcast;
    float q = 0.f;
    for (int t = tid; t < n; t += blockDim.x) {
      const float d = row[t] - mean;
      q += d * d;
    }
    q = warp_sum(q);
    __syncthreads();
    if (lane == 0) red[warp] = q;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < 8; ++w) tot += red[w];
      float sd = sqrtf(tot / ((float)n - 1.0f));  // unbiased (features.py:83-86)
      if (isnan(sd)) sd = 0.f;                     // features.py:87
      bcast = 1.0f / (sd + 1e-5f);                 // features.py:89
    }
    __syncthreads();
    inv_std = bcast;
  }
  for (int t = tid; t < T; t += blockDim.x) {
    float v = pad_value;
    if (t < n) v = normalize ? (row[t] - mean) * inv_std : row[t];
    row[t] = v;
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_mel_seq_len(const int64_t* length, int64_t* seq_len, int B, int n_fft, int hop, void* stream) {
  B2C_REQUIRE(B > 0 && hop > 0, "mel_seq_len: bad args");
  mel_seq_len_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(length, seq_len, B, n_fft, hop);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_mel_fb_ranges(const float* fb, int n_mels, int n_bins, int32_t* ranges, void* stream) {
  mel_fb_ranges_kernel<<<(n_mels + 63) / 64, 64, 0, (cudaStream_t)stream>>>(fb, n_mels, n_bins,
                                                                            reinterpret_cast<int2*>(ranges));
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_mel_frontend(const float* x, const int64_t* length, const float* window, const float* fb,
                                const int32_t* fb_ranges, float* mel, int64_t* seq_len, int B, int S, int T,
                                int n_mels, int win, int hop, int n_fft, float preemph, float dither,
                                unsigned long long seed, int normalize, float pad_value, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B2C_REQUIRE(n_fft == kNfft, "mel_frontend: only n_fft=512 is built (got %d)", n_fft);
  B2C_REQUIRE(win <= n_fft && win > 0 && hop > 0, "mel_frontend: bad window/hop");
  B2C_REQUIRE(T == S / hop + 1, "mel_frontend: T must be S/hop+1 (centre=True): S=%d hop=%d T=%d", S, hop, T);
  B2C_REQUIRE(n_mels > 0 && n_mels <= 128, "mel_frontend: n_mels out of range");
  int rc = b2c_mel_seq_len(length, seq_len, B, n_fft, hop, stream_);
  if (rc) return rc;
  const int ylen = (kFramesPerCta - 1) * hop + win;
  const size_t smem = sizeof(float) * (size_t)(((ylen + 3) & ~3) + 2 * kNfft + ((win + 3) & ~3) +
                                               n_mels * (kFramesPerCta + 1) + kMelWarps * 2 * kPad);
  B2C_REQUIRE(smem <= 227 * 1024, "mel_frontend: smem %zu too large", smem);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    B2C_CUDA(cudaFuncSetAttribute(mel_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2C_CUDA(cudaFuncSetAttribute(mel_frames_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    smem_set = smem;
  }
  dim3 grid((T + kFramesPerCta - 1) / kFramesPerCta, B);
  mel_frames_kernel<<<grid, kMelWarps * 32, smem, stream>>>(x, length, window, fb, reinterpret_cast<const int2*>(fb_ranges),
                                                           mel, S, T, n_mels, win, hop, preemph, dither, seed);
  B2C_LAUNCH_CHECK();
  mel_normalize_kernel<<<dim3(n_mels, B), 256, 0, stream>>>(mel, seq_len, T, normalize, pad_value);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
