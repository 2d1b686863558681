// tcgen05 / TMEM / TMA GEMM core for the Conformer hot path (sm_100a).
//
//   D[M,N] = sum_k A(m,k) * B(n,k)        bf16 operands, fp32 accumulation in TMEM
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer  (cp.async.bulk.tensor → 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      MMA issuer    (one thread: tcgen05.mma cta_group::1, M=128, N=BN, K=16 per instruction)
//   warp 2      TMEM allocator
//   warps 4..7  epilogue      (tcgen05.ld 32x32b: thread = accumulator row; fused bias / SiLU / GLU / residual ...)
// Two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Operand majors (no transposes are ever materialised):
//   K-major  : global [rows, K] with K contiguous   (activations x, weights W[N,K] in forward)
//   MN-major : global [K, rows] with rows contiguous (dY / x in wgrad: contraction over tokens; W in dgrad)
//
// Replaces every nn.Linear / 1x1 Conv1d of the reference hot path:
//   conformer_modules.py:382-387 (FFN), multi_head_attention.py:104-123,300,146 (q/k/v/pos/out),
//   conformer_modules.py:322,347 (pointwise convs), subsampling.py:431 (pre_encode.out) and their autograd.
#include <stdarg.h>

#include "common.cuh"

namespace b2c {

B2C_DEFINE_SEED_SETTER(set_seed_counter_gemm)


constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kEpiWarps = 16;
constexpr int kCluster = 2;  // CTA pair = one cta_group::2 MMA: 256 x BN tile, each CTA holds its 128 rows of A and BN/2 of B
constexpr int kNumThreads = 128 + 32 * kEpiWarps;  // 4 role warps + 16 epilogue warps

// Implicit-GEMM geometry of the 3x3 stride-2 subsampling convolution (no im2col is ever materialised: the
// A tiles are gathered by 4-D TMA boxes, strided for the forward / weight-gradient, shifted for the data-gradient).
struct ConvGeom {
  int C;                // channels per tap (K per tap)
  int Fq, TR, Tq;       // f positions per time row, time rows per tile (M tile, or K block for wgrad), time rows
  int tiles_per_batch;  // ceil(Tq / TR)
  int a_t_mul;          // A time coordinate = t0 * a_t_mul + tap_t
  int ntaps;
  int tap_t[9], tap_f[9], tap_k[9];  // A coordinate offsets and B k-offset (elements) per tap
  int To, Fo, ts, fs, pt, pf;        // output row = ((b*To + t*ts + pt) * Fo + f*fs + pf)
};

struct GemmParams {
  int M, N, K;
  int splits;       // split-K factor (fp32 atomic accumulate epilogue only); conv wgrad: tap index
  int k_per_split;  // in units of BK blocks
  int tma_store;    // bf16 outputs leave through TMA stores (tensor maps tmC0 / tmC1) instead of per-thread stores
  int tma_aux;      // the bf16 side input of the SiLU' epilogue arrives through TMA loads (tensor map tmAux)
  b2c_epilogue epi;
  ConvGeom cv;
};

template <int BN, bool AUX = false>
struct SmemLayout {
  // AUX: the epilogue warps also stage a side input (2 x 1 KB per warp) — one pipeline stage fewer pays for it
  static constexpr int kStages = (BN == 256) ? (AUX ? 5 : 6) : (AUX ? 6 : 8);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / kCluster) * BK * 2;  // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kAuxBar = kBarOffset + 256;                   // 2 mbarriers per epilogue warp (side-input loads)
  static constexpr int kStageOut = kBarOffset + 512;                 // per-epilogue-warp 2 KB output staging
  static constexpr int kStageAux = kStageOut + kEpiWarps * 2048;     // per-epilogue-warp 2 x 1 KB side-input staging
  static constexpr int kTotal = kStageAux + (AUX ? kEpiWarps * 2048 : 0) + 1024;  // + alignment slack
};

// ---------------------------------------------------------------- epilogue
// 16 epilogue warps (4 per TMEM lane quarter); each step handles a 32-row x EW-column chunk: thread = accumulator row.
// Outputs are bounced through a small swizzled per-warp smem tile so every global store writes whole 32 B sectors of
// a row (a thread-per-row store would touch 32 rows per instruction).  Ragged N falls back to per-thread stores.
constexpr int EW = 16;

// idx0 % 16 == 0 (the host requires N % 16 == 0 when dropout is on): one chunk hash, then a multiply-add per pair; the
// 16-bit lanes of the hash word are compared in place against the threshold shifted to the upper half
__device__ __forceinline__ void dropout_ew(float (&v)[EW], const b2c_epilogue& e, uint64_t idx0) {
  static_assert(EW == 16, "one dropout chunk per epilogue chunk");
  const uint32_t base = dropout_base(e.seed, e.drop_site, idx0 >> 4);
  const uint32_t keep_hi = dropout_keep_hi(e.drop_keep);
#pragma unroll
  for (int j = 0; j < EW; j += 2) {
    uint32_t w0, w1;
    dropout_pair(base, j >> 1, w0, w1);
    v[j] = (w0 < keep_hi) ? v[j] * e.drop_scale : 0.f;
    v[j + 1] = (w1 < keep_hi) ? v[j + 1] * e.drop_scale : 0.f;
  }
}
// round a pair to bf16 and back: one packed conversion + two bit operations (instead of four scalar conversions);
// returns the packed word (low half = a)
__device__ __forceinline__ uint32_t round_pair(float& a, float& b) {
  const uint32_t w = pack_bf16(a, b);
  a = __uint_as_float(w << 16);
  b = __uint_as_float(w & 0xFFFF0000u);
  return w;
}
__device__ __forceinline__ void round_ew(float (&v)[EW]) {
#pragma unroll
  for (int j = 0; j < EW; j += 2) round_pair(v[j], v[j + 1]);
}
__device__ __forceinline__ void load_bf16_ew(const bf16* p, float (&v)[EW]) {
#pragma unroll
  for (int j = 0; j < EW; j += 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(p + j);
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    v[j] = a.x; v[j + 1] = a.y; v[j + 2] = b.x; v[j + 3] = b.y; v[j + 4] = c.x; v[j + 5] = c.y; v[j + 6] = d.x; v[j + 7] = d.y;
  }
}
// bf16 chunk (32 rows x 16 cols = 32 B per row) → smem (swizzled) → global, 16 rows x 32 B per store instruction
__device__ __forceinline__ void stage_store_bf16(uint8_t* stg, int lane, const float (&v)[EW], bf16* out, size_t ldo,
                                                 int m_lane, int M, int n0) {
  __syncwarp();
  const uint32_t sb = smem_u32(stg);
  const int f = (lane >> 2) & 1;
#pragma unroll
  for (int k = 0; k < 2; ++k)
    sts128(sb + lane * 32 + ((k ^ f) << 4),
           make_uint4(pack_bf16(v[8 * k], v[8 * k + 1]), pack_bf16(v[8 * k + 2], v[8 * k + 3]),
                      pack_bf16(v[8 * k + 4], v[8 * k + 5]), pack_bf16(v[8 * k + 6], v[8 * k + 7])));
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int row = it * 16 + (lane >> 1), ck = lane & 1;
    const int m = __shfl_sync(0xffffffffu, m_lane, row);
    const uint4 q = lds128(sb + row * 32 + ((ck ^ ((row >> 2) & 1)) << 4));
    if (m < M) *reinterpret_cast<uint4*>(out + (size_t)m * ldo + n0 + ck * 8) = q;
  }
}
// same chunk through a TMA store: the 32 x 16 tile is written in the 32-byte-swizzle layout the tensor map expects
// (16-byte unit k of row r at r*32 + ((k ^ ((r >> 2) & 1)) << 4)), one lane issues the bulk store of the box whose first
// row is m_warp; rows >= M and columns >= N are clipped by the tensor map.  The caller has made sure `buf` is free.
__device__ __forceinline__ void stage_tma_words(uint8_t* buf, int lane, const uint32_t (&w)[EW / 2], const CUtensorMap* tm,
                                                int m_warp, int n0) {
  const uint32_t sb = smem_u32(buf);
  const int f = (lane >> 2) & 1;
#pragma unroll
  for (int k = 0; k < 2; ++k) sts128(sb + lane * 32 + ((k ^ f) << 4), make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]));
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) tma_store_2d(tm, buf, n0, m_warp);
}
__device__ __forceinline__ void stage_tma_bf16(uint8_t* buf, int lane, const float (&v)[EW], const CUtensorMap* tm, int m_warp,
                                               int n0) {
  uint32_t w[EW / 2];
#pragma unroll
  for (int j = 0; j < EW / 2; ++j) w[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
  stage_tma_words(buf, lane, w, tm, m_warp, n0);
}
// fp32 chunk (32 rows x 16 cols = 64 B per row): out = residual + v, 8 rows x 64 B per store instruction.
// The residual (aux) tile is fetched by resid_prefetch() BEFORE the epilogue maths so its latency is hidden.
struct ResidTile {
  float4 r[4];
  int mm[4];
};
__device__ __forceinline__ void resid_prefetch(ResidTile& t, const float* aux, size_t ldo, int m_lane, int M, int n0, int lane) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int row = it * 8 + (lane >> 2), ck = lane & 3;
    t.mm[it] = __shfl_sync(0xffffffffu, m_lane, row);
    t.r[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (aux && t.mm[it] < M) t.r[it] = __ldg(reinterpret_cast<const float4*>(aux + (size_t)t.mm[it] * ldo + n0 + ck * 4));
  }
}
__device__ __forceinline__ void stage_store_f32(uint8_t* stg, int lane, const float (&v)[EW], float* out, const ResidTile& t,
                                                size_t ldo, int M, int n0) {
  __syncwarp();
  const uint32_t sb = smem_u32(stg);
  const int f = (lane >> 1) & 3;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    sts128(sb + lane * 64 + ((k ^ f) << 4), make_uint4(__float_as_uint(v[4 * k]), __float_as_uint(v[4 * k + 1]),
                                                       __float_as_uint(v[4 * k + 2]), __float_as_uint(v[4 * k + 3])));
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int row = it * 8 + (lane >> 2), ck = lane & 3;
    const uint4 q = lds128(sb + row * 64 + ((ck ^ ((row >> 1) & 3)) << 4));
    if (t.mm[it] < M)
      *reinterpret_cast<float4*>(out + (size_t)t.mm[it] * ldo + n0 + ck * 4) =
          make_float4(__uint_as_float(q.x) + t.r[it].x, __uint_as_float(q.y) + t.r[it].y, __uint_as_float(q.z) + t.r[it].z,
                      __uint_as_float(q.w) + t.r[it].w);
  }
}

// per-thread fallback for ragged N (n0 + EW > N) and for the fp32 accumulate epilogue
template <int MODE>
__device__ __forceinline__ void epilogue_row(const b2c_epilogue& e, int m, int n0, int M, int N, float (&v)[EW],
                                             bool atomic) {
  if (m >= M || n0 >= N) return;
  const size_t ldo = (size_t)e.ldo;
  const uint64_t didx = (uint64_t)m * (uint64_t)N + (uint64_t)n0;
  bool valid = true;
  if ((MODE == B2C_EPI_GLU_MASK || MODE == B2C_EPI_RELU_MASK) && e.lengths) {
    const int b = m / e.rows_per_batch;
    const int t = (MODE == B2C_EPI_GLU_MASK) ? (m - b * e.rows_per_batch) : (m - b * e.rows_per_batch) / e.rows_inner;
    valid = t < e.lengths[b];
  }
  if (MODE != B2C_EPI_ACCUM_F32 && MODE != B2C_EPI_RELU_BWD && MODE != B2C_EPI_DSILU && e.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < EW; ++j) if (n0 + j < N) v[j] += e.bias[n0 + j];
  }
  if (MODE == B2C_EPI_ACCUM_F32) {
    float* o = reinterpret_cast<float*>(e.out0) + (size_t)m * ldo + n0;
    const bool vec_ok = (n0 + EW <= N) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
    if (atomic) {
      if (vec_ok) {
#pragma unroll
        for (int j = 0; j < EW; j += 4) red_add_v4(o + j, e.alpha * v[j], e.alpha * v[j + 1], e.alpha * v[j + 2], e.alpha * v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < EW; ++j) if (n0 + j < N) atomicAdd(o + j, e.alpha * v[j]);
      }
    } else if (vec_ok) {
#pragma unroll
      for (int j = 0; j < EW; j += 4) {
        float4 c = *reinterpret_cast<float4*>(o + j);
        c.x = fmaf(e.alpha, v[j], c.x); c.y = fmaf(e.alpha, v[j + 1], c.y);
        c.z = fmaf(e.alpha, v[j + 2], c.z); c.w = fmaf(e.alpha, v[j + 3], c.w);
        *reinterpret_cast<float4*>(o + j) = c;
      }
    } else {
#pragma unroll
      for (int j = 0; j < EW; ++j) if (n0 + j < N) o[j] += e.alpha * v[j];
    }
    return;
  }
  // ragged-N path of the other modes: straightforward per-element code (rare: N % 16 != 0)
#pragma unroll
  for (int j = 0; j < EW; ++j) {
    const int n = n0 + j;
    if (n >= N) continue;
    float x = v[j];
    const bool keep = (e.drop_keep == 0) || dropout_keep(e.seed, e.drop_site, didx + j, e.drop_keep);
    if (MODE == B2C_EPI_BF16) {
      reinterpret_cast<bf16*>(e.out0)[(size_t)m * ldo + n] = __float2bfloat16_rn(x);
    } else if (MODE == B2C_EPI_SILU) {
      float h = silu_f(x);
      if (e.drop_keep != 0) h = keep ? h * e.drop_scale : 0.f;
      reinterpret_cast<bf16*>(e.out0)[(size_t)m * ldo + n] = __float2bfloat16_rn(x);
      reinterpret_cast<bf16*>(e.out1)[(size_t)m * ldo + n] = __float2bfloat16_rn(h);
    } else if (MODE == B2C_EPI_RESID_F32) {
      x = bf16_round(x);
      if (e.drop_keep != 0) x = keep ? x * e.drop_scale : 0.f;
      const float r = e.aux ? reinterpret_cast<const float*>(e.aux)[(size_t)m * ldo + n] : 0.f;
      reinterpret_cast<float*>(e.out0)[(size_t)m * ldo + n] = r + e.alpha * x;
    } else if (MODE == B2C_EPI_GLU_MASK) {
      x = bf16_round(x);
      reinterpret_cast<bf16*>(e.out0)[(size_t)m * ldo + n] = __float2bfloat16_rn(x);
      if (j & 1) {
        const float a = bf16_round(v[j - 1]);  // value column (bias already added above)
        reinterpret_cast<bf16*>(e.out1)[(size_t)m * (ldo >> 1) + (n >> 1)] = __float2bfloat16_rn(valid ? a * sigmoid_f(x) : 0.f);
      }
    } else if (MODE == B2C_EPI_DSILU) {
      if (e.drop_keep != 0) x = keep ? x * e.drop_scale : 0.f;
      const float z = __bfloat162float(reinterpret_cast<const bf16*>(e.aux)[(size_t)m * ldo + n]);
      reinterpret_cast<bf16*>(e.out0)[(size_t)m * ldo + n] = __float2bfloat16_rn(x * dsilu_f(z));
    } else if (MODE == B2C_EPI_RELU_MASK) {
      reinterpret_cast<bf16*>(e.out0)[(size_t)m * ldo + n] = __float2bfloat16_rn(valid ? fmaxf(x, 0.f) : 0.f);
    } else if (MODE == B2C_EPI_RELU_BWD) {
      const float y = __bfloat162float(reinterpret_cast<const bf16*>(e.aux)[(size_t)m * ldo + n]);
      reinterpret_cast<bf16*>(e.out0)[(size_t)m * ldo + n] = __float2bfloat16_rn(y > 0.f ? x : 0.f);
    }
  }
}

// Warp-collective epilogue of one 32-row x EW-column chunk.
// bf16 side input (z for SiLU', y for ReLU') of one row chunk, fetched one chunk ahead of its use
// bf16 outputs of the plain GEMMs leave through TMA stores: per epilogue warp two 1 KB staging buffers; `turn`
// alternates them for single-output modes, the SiLU mode (two outputs) uses both for every chunk
struct TmaOut {
  const CUtensorMap* tm0;
  const CUtensorMap* tm1;
  int m_warp;  // first accumulator row of this warp's 32 rows
  int on;
  int turn;
};
template <int MODE>
constexpr bool kTmaStoreMode = (MODE == B2C_EPI_BF16 || MODE == B2C_EPI_SILU || MODE == B2C_EPI_GLU_MASK || MODE == B2C_EPI_DSILU ||
                                MODE == B2C_EPI_RELU_MASK || MODE == B2C_EPI_RELU_BWD);
// out0 chunk: TMA path (waits until the staging buffer it is about to overwrite has been read) or per-thread stores
__device__ __forceinline__ void store_out0_bf16(uint8_t* stg, int lane, const float (&v)[EW], bf16* out, size_t ldo, int m, int M,
                                                int n0, bool ts, const TmaOut& to) {
  if (ts) {
    if (lane == 0) bulk_wait_read<1>();
    __syncwarp();
    stage_tma_bf16(stg + (to.turn & 1) * 1024, lane, v, to.tm0, to.m_warp, n0);
    if (lane == 0) bulk_commit();
  } else {
    stage_store_bf16(stg, lane, v, out, ldo, m, M, n0);
  }
}
struct AuxChunk {
  uint4 q[2];
};
template <int MODE>
__device__ __forceinline__ void aux_prefetch(AuxChunk& a, const b2c_epilogue& e, int m, int n0, int M, int N) {
  a.q[0] = a.q[1] = make_uint4(0, 0, 0, 0);
  if ((MODE == B2C_EPI_DSILU || MODE == B2C_EPI_RELU_BWD) && m < M && n0 + EW <= N) {
    const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(e.aux) + (size_t)m * (size_t)e.ldo + n0);
    a.q[0] = __ldg(p);
    a.q[1] = __ldg(p + 1);
  }
}
__device__ __forceinline__ void unpack_aux(const AuxChunk& a, float (&v)[EW]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float2 x0 = unpack_bf16(a.q[h].x), x1 = unpack_bf16(a.q[h].y), x2 = unpack_bf16(a.q[h].z), x3 = unpack_bf16(a.q[h].w);
    v[8 * h] = x0.x; v[8 * h + 1] = x0.y; v[8 * h + 2] = x1.x; v[8 * h + 3] = x1.y;
    v[8 * h + 4] = x2.x; v[8 * h + 5] = x2.y; v[8 * h + 6] = x3.x; v[8 * h + 7] = x3.y;
  }
}

template <int MODE>
__device__ __forceinline__ void epilogue_chunk(const b2c_epilogue& e, int m, int n0, int M, int N, const uint32_t (&acc)[EW],
                                               const AuxChunk& auxc, bool atomic, uint8_t* stg, int lane, const TmaOut& to) {
  if (n0 >= N) return;  // warp-uniform
  float v[EW];
#pragma unroll
  for (int j = 0; j < EW; ++j) v[j] = __uint_as_float(acc[j]);
  // TMA-stored outputs: ragged rows / columns are clipped by the tensor map, no per-thread fallback needed
  const bool ts = kTmaStoreMode<MODE> && to.on;
  if (MODE == B2C_EPI_ACCUM_F32 || ((!ts || MODE == B2C_EPI_GLU_MASK) && n0 + EW > N)) {
    epilogue_row<MODE>(e, m, n0, M, N, v, atomic);
    return;
  }
  const bool rowok = m < M;
  const int ms = rowok ? m : 0;  // safe row for the loads of inactive lanes
  ResidTile rt;
  if (MODE == B2C_EPI_RESID_F32) resid_prefetch(rt, reinterpret_cast<const float*>(e.aux), (size_t)e.ldo, m, M, n0, lane);
  if (MODE != B2C_EPI_RELU_BWD && MODE != B2C_EPI_DSILU && e.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < EW; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + j));
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  const size_t ldo = (size_t)e.ldo;
  const uint64_t didx = (uint64_t)ms * (uint64_t)N + (uint64_t)n0;
  bool valid = rowok;
  if ((MODE == B2C_EPI_GLU_MASK || MODE == B2C_EPI_RELU_MASK) && e.lengths && rowok) {
    const int b = m / e.rows_per_batch;
    const int t = (MODE == B2C_EPI_GLU_MASK) ? (m - b * e.rows_per_batch) : (m - b * e.rows_per_batch) / e.rows_inner;
    valid = t < e.lengths[b];
  }
  if (MODE == B2C_EPI_BF16) {
    store_out0_bf16(stg, lane, v, reinterpret_cast<bf16*>(e.out0), ldo, m, M, n0, ts, to);
  } else if (MODE == B2C_EPI_SILU) {
    // z = bf16(acc+bias) → out0 ; h = bf16(dropout(silu(acc+bias))) → out1   (conformer_modules.py:383-385).
    // silu and the dropout scale act on the fp32 value and h is rounded once (the reference under autocast rounds z, silu(z)
    // and the dropout product to bf16 one after the other): one rounding instead of three, and ~40 fewer instructions per
    // 16-element chunk in an epilogue that is bound by instruction issue (K = 512: a 128x256 tile every 2.2 us of MMA time)
    float h[EW];
    uint32_t zw[EW / 2];
#pragma unroll
    for (int j = 0; j < EW; j += 2) {
      zw[j >> 1] = pack_bf16(v[j], v[j + 1]);
      h[j] = silu_f(v[j]);
      h[j + 1] = silu_f(v[j + 1]);
    }
    if (e.drop_keep != 0) dropout_ew(h, e, didx);
    if (ts) {
      uint32_t hw[EW / 2];
#pragma unroll
      for (int j = 0; j < EW / 2; ++j) hw[j] = pack_bf16(h[2 * j], h[2 * j + 1]);
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
      stage_tma_words(stg, lane, zw, to.tm0, to.m_warp, n0);
      stage_tma_words(stg + 1024, lane, hw, to.tm1, to.m_warp, n0);
      if (lane == 0) bulk_commit();
    } else {
      stage_store_bf16(stg, lane, v, reinterpret_cast<bf16*>(e.out0), ldo, m, M, n0);
      stage_store_bf16(stg, lane, h, reinterpret_cast<bf16*>(e.out1), ldo, m, M, n0);
    }
  } else if (MODE == B2C_EPI_RESID_F32) {
    // out0_f32 = (aux_f32 ? aux : 0) + alpha * dropout(bf16(acc+bias))
    round_ew(v);
    if (e.drop_keep != 0) dropout_ew(v, e, didx);
#pragma unroll
    for (int j = 0; j < EW; ++j) v[j] *= e.alpha;
    stage_store_f32(stg, lane, v, reinterpret_cast<float*>(e.out0), rt, ldo, M, n0);
  } else if (MODE == B2C_EPI_GLU_MASK) {
    // interleaved (value, gate): g = bf16(acc+bias) → out0 [M,N]; u = bf16(a*sigmoid(g))*(t<len) → out1 [M,N/2]
    float u[EW / 2];
#pragma unroll
    for (int j = 0; j < EW; j += 2) {
      round_pair(v[j], v[j + 1]);
      u[j >> 1] = valid ? v[j] * sigmoid_f(v[j + 1]) : 0.f;
    }
    store_out0_bf16(stg, lane, v, reinterpret_cast<bf16*>(e.out0), ldo, m, M, n0, ts, to);
    if (rowok)
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(e.out1) + (size_t)m * (ldo >> 1) + (n0 >> 1)) =
          make_uint4(pack_bf16(u[0], u[1]), pack_bf16(u[2], u[3]), pack_bf16(u[4], u[5]), pack_bf16(u[6], u[7]));
  } else if (MODE == B2C_EPI_DSILU) {
    // dz = bf16(dropout(bf16(acc)) * silu'(z)),  z = aux (bf16)
    float zz[EW];
    unpack_aux(auxc, zz);
    if (e.drop_keep != 0) dropout_ew(v, e, didx);  // one rounding, at the store (see the SiLU epilogue)
#pragma unroll
    for (int j = 0; j < EW; ++j) v[j] *= dsilu_f(zz[j]);
    store_out0_bf16(stg, lane, v, reinterpret_cast<bf16*>(e.out0), ldo, m, M, n0, ts, to);
  } else if (MODE == B2C_EPI_RELU_MASK) {
#pragma unroll
    for (int j = 0; j < EW; ++j) v[j] = valid ? fmaxf(v[j], 0.f) : 0.f;
    store_out0_bf16(stg, lane, v, reinterpret_cast<bf16*>(e.out0), ldo, m, M, n0, ts, to);
  } else if (MODE == B2C_EPI_RELU_BWD) {
    float yy[EW];
    unpack_aux(auxc, yy);
#pragma unroll
    for (int j = 0; j < EW; ++j) v[j] = yy[j] > 0.f ? v[j] : 0.f;
    store_out0_bf16(stg, lane, v, reinterpret_cast<bf16*>(e.out0), ldo, m, M, n0, ts, to);
  }
}

// ---------------------------------------------------------------- kernel
// CONV: 0 = plain GEMM; 1 = conv forward (strided gather A); 2 = conv data-gradient (shifted gather A, scatter rows);
//       3 = conv weight-gradient (tokens are the contraction; A = dY, B = strided gather of X, one tap per `split`)
template <int BN, int A_MN, int B_MN, int MODE, int CONV = 0>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC0, const __grid_constant__ CUtensorMap tmC1,
               const __grid_constant__ CUtensorMap tmAux, const GemmParams p) {
  constexpr bool kAux = (MODE == B2C_EPI_DSILU && CONV == 0);
  using L = SmemLayout<BN, kAux>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (CONV == 1 || CONV == 2) ? p.M : (p.M + BM - 1) / BM;  // conv fwd/dgrad: p.M = #row tiles
  const int n_tiles = (p.N + BN - 1) / BN;
  // Tiles are enumerated so that the two CTAs of a cluster (blockIdx.x = 2c, 2c+1; persistent stride is even) always
  // work on M-adjacent tiles of the SAME n-tile / split at the same time: together they form one 256 x BN cta_group::2
  // tile.  Each CTA TMA-loads its own 128 rows of A and its half of the B tile into its own shared memory, the leader
  // (even) CTA issues the MMAs for both, each CTA's TMEM receives its 128 accumulator rows and each runs its own
  // epilogue.  An odd tile row count adds one dummy row tile (TMA zero-fills, epilogue skips).
  const int m_pairs = (m_tiles + kCluster - 1) / kCluster;
  const int total_tiles = m_pairs * kCluster * n_tiles * p.splits;
  const uint32_t crank = (kCluster > 1) ? cluster_ctarank() : 0u;
  auto decode = [&](int tile, int& m_blk, int& n_blk, int& split) {
    const int pair = tile / kCluster;
    split = pair % p.splits;
    const int mn2 = pair / p.splits;
    n_blk = mn2 % n_tiles;
    m_blk = (mn2 / n_tiles) * kCluster + (tile % kCluster);
  };
  const int kb_total = (CONV != 0) ? p.K : (p.K + BK - 1) / BK;               // conv: p.K = #k-blocks
  // bytes one stage receives (partial boxes of the conv gathers deliver fewer rows than the slot holds)
  const uint32_t stage_tx = (CONV == 1 || CONV == 2) ? (uint32_t)(p.cv.TR * p.cv.Fq * BK * 2 + L::kBBytes)
                            : (CONV == 3)            ? (uint32_t)((BM / 64 + BN / 64 / kCluster) * p.cv.TR * p.cv.Fq * 128)
                                                     : (uint32_t)L::kStageBytes;
  if (CONV == 3) {
    // rows the partial boxes never write must read as zero for the MN-major contraction
    for (int i = threadIdx.x; i < kStages * L::kStageBytes / 16; i += kNumThreads)
      reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);   // leader only: its own expect_tx; both CTAs' TMA bytes complete on it
      mbar_init(&empty_bar[i], 1);  // the leader's commit arrives on both CTAs' barriers
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kCluster * kEpiWarps);  // leader only: one arrive per epilogue warp of BOTH CTAs
    }
    if (kAux)
      for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(reinterpret_cast<uint64_t*>(smem + L::kAuxBar) + i, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers are initialised and its TMEM allocated before any remote arrive / paired MMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_blk, n_blk, split;
        decode(tile, m_blk, n_blk, split);
        const int m0 = m_blk * BM, n0 = n_blk * BN + (int)crank * (BN / kCluster);  // this CTA's half of the B tile
        // conv weight-gradient: split = k_split * 9 + tap
        const int kb0 = (CONV == 3) ? (split / 9) * p.k_per_split : split * p.k_per_split;
        const int kb1 = min(kb_total, kb0 + p.k_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          uint64_t* fb = &full_bar[stage];  // every load below signals the LEADER's barrier at this offset
          if (crank == 0) mbar_expect_tx(fb, kCluster * stage_tx);
          // ---- A operand: this CTA's 128 rows
          if (CONV == 1 || CONV == 2) {
            const int cblks = p.cv.C / BK;
            const int tap = kb / cblks, c0 = (kb - tap * cblks) * BK;
            const int bb = m_blk / p.cv.tiles_per_batch, t0 = (m_blk - bb * p.cv.tiles_per_batch) * p.cv.TR;
            tma_load_4d_2sm(sa, &tmA, fb, c0, p.cv.tap_f[tap], t0 * p.cv.a_t_mul + p.cv.tap_t[tap], bb);
          } else if (CONV == 3) {
            const int bb = kb / p.cv.tiles_per_batch, t0 = (kb - bb * p.cv.tiles_per_batch) * p.cv.TR;
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_4d_2sm(sa + c * (64 * BK * 2), &tmA, fb, m0 + c * 64, 0, t0, bb);
          } else if (A_MN) {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d_2sm(sa + c * (64 * BK * 2), &tmA, fb, m0 + c * 64, kb * BK);
          } else {
            tma_load_2d_2sm(sa, &tmA, fb, kb * BK, m0);
          }
          // ---- B operand: this CTA's BN/2 rows (K-major) or columns (MN-major) of the tile
          if (CONV == 3) {
            const int bb = kb / p.cv.tiles_per_batch, t0 = (kb - bb * p.cv.tiles_per_batch) * p.cv.TR;
            const int tap = split % 9;  // kh*3 + kw
#pragma unroll
            for (int c = 0; c < BN / 64 / kCluster; ++c)
              tma_load_4d_2sm(sb + c * (64 * BK * 2), &tmB, fb, n0 + c * 64, tap % 3 - 1, 2 * t0 + tap / 3 - 1, bb);
          } else if (CONV == 1 || CONV == 2) {
            const int cblks = p.cv.C / BK;
            const int tap = kb / cblks, c0 = (kb - tap * cblks) * BK;
            tma_load_2d_2sm(sb, &tmB, fb, p.cv.tap_k[tap] + c0, n0);
          } else if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64 / kCluster; ++c) tma_load_2d_2sm(sb + c * (64 * BK * 2), &tmB, fb, n0 + c * 64, kb * BK);
          } else {
            tma_load_2d_2sm(sb, &tmB, fb, kb * BK, n0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {  // the leader CTA issues the pair's MMAs
      constexpr uint32_t idesc = umma_idesc_bf16(kCluster * BM, BN, A_MN, B_MN);
      constexpr uint16_t kBoth = (uint16_t)((1u << kCluster) - 1);
      int stage = 0, phase = 0, as = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int m_blk, n_blk, split;
        decode(tile, m_blk, n_blk, split);
        // conv weight-gradient: split = k_split * 9 + tap
        const int kb0 = (CONV == 3) ? (split / 9) * p.k_per_split : split * p.k_per_split;
        const int kb1 = min(kb_total, kb0 + p.k_per_split);
        mbar_wait(&tempty_bar[as], aphase ^ 1);  // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);    // both CTAs' operand tiles have landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: +32 B per K=16 inside the 128 B swizzle row; SBO = 8 rows * 128 B.
            // MN-major: 16 k-rows * 128 B = 2048 B per K=16; LBO = stride between 64-wide MN chunks (64 k-rows*128 B).
            // The same descriptors (= shared-memory offsets) are applied in both CTAs.
            const uint64_t da = A_MN ? umma_desc_sw128(sa + k * 2048, 64 * BK * 2, 1024) : umma_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_desc_sw128(sb + k * 2048, 64 * BK * 2, 1024) : umma_desc_sw128(sb + k * 32, 16, 1024);
            umma_bf16_2sm(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage], kBoth);  // frees the smem slot in BOTH CTAs once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&tfull_bar[as], kBoth);  // accumulator complete → both epilogues
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;             // TMEM lane quarter this warp may access (= warp id % 4)
    const int cgrp = (warp - 4) >> 2;   // four warps share a quarter and split the column chunks
    constexpr int kChunks0 = BN / EW / (kEpiWarps / 4);
    uint8_t* stg = smem + L::kStageOut + (warp - 4) * 2048;
    TmaOut to;
    to.tm0 = &tmC0; to.tm1 = &tmC1; to.on = (CONV == 0) ? p.tma_store : 0; to.turn = 0; to.m_warp = 0;
    // side input through TMA (SiLU' epilogue): 32-row x 16-column boxes into two 1 KB buffers, one mbarrier each
    const bool aux_tma = kAux && p.tma_aux != 0;
    uint8_t* aux_buf = smem + L::kStageAux + (warp - 4) * 2048;
    uint64_t* aux_bar = reinterpret_cast<uint64_t*>(smem + L::kAuxBar) + (warp - 4) * 2;
    uint32_t aux_use = 0;
    if (aux_tma && lane == 0 && (int)blockIdx.x < total_tiles) {  // first chunk of the first tile
      int mb0, nb0, sp0;
      decode(blockIdx.x, mb0, nb0, sp0);
      mbar_expect_tx(&aux_bar[0], 1024);
      tma_load_2d(aux_buf, &tmAux, &aux_bar[0], nb0 * BN + cgrp * kChunks0 * EW, mb0 * BM + q * 32);
    }
    int as = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int m_blk, n_blk, split;
      decode(tile, m_blk, n_blk, split);
      const int m0 = m_blk * BM, n0 = n_blk * BN;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      int m = m0 + q * 32 + lane, m_bound = p.M;
      b2c_epilogue epi = p.epi;
      if (CONV == 1 || CONV == 2) {
        const int mb = m_blk;
        const int bb = mb / p.cv.tiles_per_batch, t0 = (mb - bb * p.cv.tiles_per_batch) * p.cv.TR;
        const int r = q * 32 + lane;
        const int tl = r / p.cv.Fq, fl = r - tl * p.cv.Fq, t = t0 + tl;
        m_bound = 0x7fffffff;
        m = (mb < m_tiles && r < p.cv.TR * p.cv.Fq && t < p.cv.Tq)
                ? ((bb * p.cv.To + t * p.cv.ts + p.cv.pt) * p.cv.Fo + fl * p.cv.fs + p.cv.pf)
                : m_bound;
      } else if (CONV == 3) {
        epi.out0 = reinterpret_cast<float*>(epi.out0) + (size_t)(split % 9) * p.cv.C;  // tap-major columns
      }
      // software pipeline over this warp's column chunks: the TMEM load and the bf16 side input of chunk i+1 are in
      // flight while chunk i goes through the epilogue maths and the staged store
      constexpr int kChunks = BN / EW / (kEpiWarps / 4);
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      const bool atomic = (CONV == 0 && p.splits > 1) || (CONV == 3 && p.splits > 9);
      uint32_t acc[2][EW];
      AuxChunk auxc[2];
      to.m_warp = m0 + q * 32;
      // each warp owns kChunks CONSECUTIVE 16-column chunks (adjacent chunks of a row leave the SM back to back)
      tmem_ld16(trow + (uint32_t)(cgrp * kChunks * EW), acc[0]);
      if (!aux_tma) aux_prefetch<MODE>(auxc[0], epi, m, n0 + cgrp * kChunks * EW, m_bound, p.N);
#pragma unroll
      for (int i = 0; i < kChunks; ++i) {
        const int c = cgrp * kChunks + i;
        if (kAux && aux_tma) {
          // side input of chunk i+1 (or of the next tile's first chunk) → the staging buffer chunk i-1 was read from
          int nm = to.m_warp, nn = n0 + (c + 1) * EW;
          bool more = true;
          if (i + 1 == kChunks) {
            const int nt = tile + gridDim.x;
            more = nt < total_tiles;
            if (more) {
              int mb2, nb2, sp2;
              decode(nt, mb2, nb2, sp2);
              nm = mb2 * BM + q * 32;
              nn = nb2 * BN + cgrp * kChunks * EW;
            }
          }
          __syncwarp();
          if (more && lane == 0) {
            const int nb = (aux_use + 1) & 1;
            mbar_expect_tx(&aux_bar[nb], 1024);
            tma_load_2d(aux_buf + nb * 1024, &tmAux, &aux_bar[nb], nn, nm);
          }
          mbar_wait(&aux_bar[aux_use & 1], (aux_use >> 1) & 1);
          const uint32_t ab = smem_u32(aux_buf + (aux_use & 1) * 1024) + lane * 32;
          const int f = (lane >> 2) & 1;
          auxc[i & 1].q[0] = lds128(ab + ((0 ^ f) << 4));
          auxc[i & 1].q[1] = lds128(ab + ((1 ^ f) << 4));
          ++aux_use;
        }
        tmem_ld_wait();
        if (i + 1 < kChunks) {
          tmem_ld16(trow + (uint32_t)((c + 1) * EW), acc[(i + 1) & 1]);
          if (!aux_tma) aux_prefetch<MODE>(auxc[(i + 1) & 1], epi, m, n0 + (c + 1) * EW, m_bound, p.N);
        }
        epilogue_chunk<MODE>(epi, m, n0 + c * EW, m_bound, p.N, acc[i & 1], auxc[i & 1], atomic, stg, lane, to);
        to.turn ^= 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (kTmaStoreMode<MODE> && to.on && lane == 0) bulk_wait_all();  // shared-memory sources stay valid until read
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer's MMAs / arrives can still touch it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm<2 * BN>(tmem_base);
  }
}

// ---------------------------------------------------------------- debug / tiny-shape SIMT GEMM (same epilogues)
// One thread per (row, 32-column chunk); used for shapes the TMA path cannot take (row pitch not 16 B aligned)
// and by the unit tests to bisect the tcgen05 path (B2C_GEMM_IMPL=simt).
template <int MODE>
__global__ void gemm_simt_kernel(const bf16* __restrict__ A, const bf16* __restrict__ B, int lda, int ldb, int a_mn,
                                 int b_mn, const GemmParams p) {
  const int chunks = (p.N + EW - 1) / EW;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)p.M * chunks) return;
  const int m = (int)(gid / chunks), n0 = (int)(gid % chunks) * EW;
  float acc[EW];
#pragma unroll
  for (int j = 0; j < EW; ++j) acc[j] = 0.f;
  for (int k = 0; k < p.K; ++k) {
    const float a = __bfloat162float(a_mn ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k]);
#pragma unroll
    for (int j = 0; j < EW; ++j) {
      const int n = n0 + j;
      if (n < p.N) acc[j] += a * __bfloat162float(b_mn ? B[(size_t)k * ldb + n] : B[(size_t)n * ldb + k]);
    }
  }
  // the per-thread path of every mode (identical arithmetic to the staged tensor-core epilogue)
  if (MODE != B2C_EPI_ACCUM_F32) {
    // force the ragged branch of epilogue_row by pretending the chunk is ragged: it handles full chunks as well
  }
  epilogue_row<MODE>(p.epi, m, n0, p.M, p.N, acc, false);
}

// ---------------------------------------------------------------- host launch
static int g_gemm_impl = -1;  // 0 = tcgen05, 1 = simt (debug)

template <int BN, int A_MN, int B_MN, int MODE, int CONV = 0>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream,
                     const CUtensorMap* tc0 = nullptr, const CUtensorMap* tc1 = nullptr, const CUtensorMap* taux = nullptr) {
  static const CUtensorMap kNoMap = {};
  const CUtensorMap& c0 = tc0 ? *tc0 : kNoMap;
  const CUtensorMap& c1 = tc1 ? *tc1 : (tc0 ? *tc0 : kNoMap);
  const CUtensorMap& cx = taux ? *taux : kNoMap;
  using L = SmemLayout<BN, (MODE == B2C_EPI_DSILU && CONV == 0)>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, MODE, CONV>;
  static bool attr_set = false;
  if (!attr_set) {
    B2C_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  const int m_tiles = (CONV == 1 || CONV == 2) ? p.M : (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
  const int total = ((m_tiles + kCluster - 1) / kCluster) * kCluster * n_tiles * p.splits;
  int grid = total < num_sms() ? total : num_sms();
  grid -= grid % kCluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  count_launch();
  B2C_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, c0, c1, cx, p));
  return B2C_OK;
}

template <int MODE>
static int dispatch_mode(const void* A, const void* B, int lda, int ldb, int a_mn, int b_mn, GemmParams p,
                         cudaStream_t stream) {
  if (g_gemm_impl < 0) {
    const char* e = getenv("B2C_GEMM_IMPL");
    g_gemm_impl = (e && !strcmp(e, "simt")) ? 1 : 0;
  }
  const bool tma_ok = (lda % 8 == 0) && (ldb % 8 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  if (g_gemm_impl == 1 || !tma_ok) {
    p.splits = 1;
    p.k_per_split = (p.K + BK - 1) / BK;
    const long long work = (long long)p.M * ((p.N + EW - 1) / EW);
    gemm_simt_kernel<MODE><<<(unsigned)((work + 127) / 128), 128, 0, stream>>>(
        reinterpret_cast<const bf16*>(A), reinterpret_cast<const bf16*>(B), lda, ldb, a_mn, b_mn, p);
    B2C_LAUNCH_CHECK();
    return B2C_OK;
  }
  // tile width: 256 when it divides the work well, else 128.  Small-M launches (2 utterances per GPU: 12 row tiles) leave
  // most CTA pairs idle with 256-wide tiles: take the width whose (rounds of CTA pairs) x (cost per tile) is smaller —
  // a 128-wide tile costs ~0.55 of a 256-wide one (half the MMAs, less operand reuse).  Split-K (fp32 accumulate) launches
  // fill the machine through their split factor instead.
  int BN = (p.N >= 256 && p.N % 256 == 0) ? 256 : 128;
  if (BN == 256 && MODE != B2C_EPI_ACCUM_F32) {
    const int clusters = num_sms() / kCluster;
    const long long t256 = (long long)((((p.M + BM - 1) / BM) + kCluster - 1) / kCluster) * (p.N / 256);
    const double c256 = (double)((t256 + clusters - 1) / clusters);
    const double c128 = 0.55 * (double)((2 * t256 + clusters - 1) / clusters);
    if (c128 < c256) BN = 128;
  }
  const int kb_total = (p.K + BK - 1) / BK;
  if (MODE == B2C_EPI_ACCUM_F32) {
    // split-K so that the (pair-tile x split) work units fill whole waves of CTA pairs: minimise
    // rounds * (k-blocks per unit + fixed per-unit cost of the atomic epilogue), fewest splits on ties
    const int pair_tiles = (((p.M + BM - 1) / BM + kCluster - 1) / kCluster) * ((p.N + BN - 1) / BN);
    const int clusters = num_sms() / kCluster;
    int best_s = 1;
    long long best_t = -1;
    for (int sp = 1; sp <= 64 && sp <= kb_total; ++sp) {
      const int kper = (kb_total + sp - 1) / sp;
      if (sp > 1 && kper < 8) break;
      const int s_eff = (kb_total + kper - 1) / kper;
      const long long units = (long long)pair_tiles * s_eff;
      const long long rounds = (units + clusters - 1) / clusters;
      const long long t = rounds * (kper + 6);
      if (best_t < 0 || t < best_t) { best_t = t; best_s = s_eff; }
    }
    p.k_per_split = (kb_total + best_s - 1) / best_s;
    p.splits = (kb_total + p.k_per_split - 1) / p.k_per_split;
  } else {
    p.splits = 1;
    p.k_per_split = kb_total;
  }
  CUtensorMap ta, tb;
  {
    // K-major: dims {K, rows}, box {64, 128 (A) | BN/kCluster (B)};  MN-major: dims {rows, K}, box {64, 64}
    uint64_t dims[2], strides[1];
    uint32_t box[2], es[2] = {1, 1};
    if (a_mn) { dims[0] = p.M; dims[1] = p.K; box[0] = 64; box[1] = BK; }
    else      { dims[0] = p.K; dims[1] = p.M; box[0] = BK; box[1] = BM; }
    strides[0] = (uint64_t)lda * 2;
    int rc = make_tmap(&ta, A, 2, 2, dims, strides, box, es, 1);
    if (rc) return rc;
    if (b_mn) { dims[0] = p.N; dims[1] = p.K; box[0] = 64; box[1] = BK; }
    else      { dims[0] = p.K; dims[1] = p.N; box[0] = BK; box[1] = BN / kCluster; }  // each CTA of the pair loads its share
    strides[0] = (uint64_t)ldb * 2;
    rc = make_tmap(&tb, B, 2, 2, dims, strides, box, es, 1);
    if (rc) return rc;
  }
  // bf16 outputs through TMA stores when the output rows are 16-byte aligned (box = one 32-row x 16-column chunk)
  CUtensorMap tc0, tc1;
  const CUtensorMap *pc0 = nullptr, *pc1 = nullptr;
  p.tma_store = 0;
  if (kTmaStoreMode<MODE> && p.epi.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(p.epi.out0) & 15) == 0 &&
      (MODE != B2C_EPI_SILU || (p.epi.out1 && (reinterpret_cast<uintptr_t>(p.epi.out1) & 15) == 0))) {
    uint64_t dims[2] = {(uint64_t)p.N, (uint64_t)p.M}, strides[1] = {(uint64_t)p.epi.ldo * 2};
    uint32_t box[2] = {(uint32_t)EW, 32}, es[2] = {1, 1};
    int rc = make_tmap(&tc0, p.epi.out0, 2, 2, dims, strides, box, es, 3);
    if (rc) return rc;
    pc0 = &tc0;
    if (MODE == B2C_EPI_SILU) {
      rc = make_tmap(&tc1, p.epi.out1, 2, 2, dims, strides, box, es, 3);
      if (rc) return rc;
      pc1 = &tc1;
    }
    p.tma_store = 1;
  }
  // SiLU' epilogue: the bf16 side input (same shape / pitch as the output) arrives through TMA loads
  CUtensorMap tax;
  const CUtensorMap* pax = nullptr;
  p.tma_aux = 0;
  if (MODE == B2C_EPI_DSILU && p.tma_store && p.epi.aux && (reinterpret_cast<uintptr_t>(p.epi.aux) & 15) == 0) {
    uint64_t dims[2] = {(uint64_t)p.N, (uint64_t)p.M}, strides[1] = {(uint64_t)p.epi.ldo * 2};
    uint32_t box[2] = {(uint32_t)EW, 32}, es[2] = {1, 1};
    int rc = make_tmap(&tax, p.epi.aux, 2, 2, dims, strides, box, es, 3);
    if (rc) return rc;
    pax = &tax;
    p.tma_aux = 1;
  }
#define B2C_GO(BN_, AM_, BM_) return launch_tc<BN_, AM_, BM_, MODE>(ta, tb, p, stream, pc0, pc1, pax)
  if (BN == 256) {
    if (!a_mn && !b_mn) B2C_GO(256, 0, 0);
    if (!a_mn && b_mn) B2C_GO(256, 0, 1);
    if (a_mn && b_mn) B2C_GO(256, 1, 1);
    if (a_mn && !b_mn) B2C_GO(256, 1, 0);
  } else {
    if (!a_mn && !b_mn) B2C_GO(128, 0, 0);
    if (!a_mn && b_mn) B2C_GO(128, 0, 1);
    if (a_mn && b_mn) B2C_GO(128, 1, 1);
    if (a_mn && !b_mn) B2C_GO(128, 1, 0);
  }
#undef B2C_GO
  return B2C_ERR_INVALID;
}

int gemm(const void* A, const void* B, int M, int N, int K, int lda, int ldb, int a_mn, int b_mn, int mode,
         const b2c_epilogue* epi, cudaStream_t stream) {
  B2C_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad shape %d %d %d", M, N, K);
  B2C_REQUIRE(epi && epi->out0, "gemm: epilogue output missing");
  B2C_REQUIRE(epi->drop_keep == 0 || N % 16 == 0, "gemm: dropout epilogues need N %% 16 == 0 (N = %d)", N);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.splits = 1; p.k_per_split = 0; p.epi = *epi;
  switch (mode) {
    case B2C_EPI_BF16: return dispatch_mode<B2C_EPI_BF16>(A, B, lda, ldb, a_mn, b_mn, p, stream);
    case B2C_EPI_SILU: return dispatch_mode<B2C_EPI_SILU>(A, B, lda, ldb, a_mn, b_mn, p, stream);
    case B2C_EPI_RESID_F32: return dispatch_mode<B2C_EPI_RESID_F32>(A, B, lda, ldb, a_mn, b_mn, p, stream);
    case B2C_EPI_GLU_MASK: return dispatch_mode<B2C_EPI_GLU_MASK>(A, B, lda, ldb, a_mn, b_mn, p, stream);
    case B2C_EPI_DSILU: return dispatch_mode<B2C_EPI_DSILU>(A, B, lda, ldb, a_mn, b_mn, p, stream);
    case B2C_EPI_RELU_MASK: return dispatch_mode<B2C_EPI_RELU_MASK>(A, B, lda, ldb, a_mn, b_mn, p, stream);
    case B2C_EPI_ACCUM_F32: return dispatch_mode<B2C_EPI_ACCUM_F32>(A, B, lda, ldb, a_mn, b_mn, p, stream);
    case B2C_EPI_RELU_BWD: return dispatch_mode<B2C_EPI_RELU_BWD>(A, B, lda, ldb, a_mn, b_mn, p, stream);
  }
  set_last_error("gemm: unknown epilogue %d", mode);
  return B2C_ERR_INVALID;
}

// ---------------------------------------------------------------- 3x3 stride-2 convolution as implicit GEMM
static int conv_map4d(CUtensorMap* m, const void* base, int C, int F, int T, int B, int box_c, int box_f, int box_t,
                      int stride_ft) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)F, (uint64_t)T, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)F * C * 2, (uint64_t)T * F * C * 2};
  uint32_t box[4] = {(uint32_t)box_c, (uint32_t)(box_f * stride_ft), (uint32_t)(box_t * stride_ft), 1};
  uint32_t es[4] = {1, (uint32_t)stride_ft, (uint32_t)stride_ft, 1};
  return make_tmap(m, base, 2, 4, dims, strides, box, es, 1);
}

int conv3x3s2_fwd(const void* in, const void* w_bf16, const float* bias, const int32_t* len_out, void* out, int B,
                  int T1, int F1, int C, cudaStream_t stream) {
  const int T2 = (T1 - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  B2C_REQUIRE(C % 64 == 0 && F2 >= 1 && F2 <= 128 && B > 0, "conv3x3s2_fwd: unsupported shape C=%d F2=%d", C, F2);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  ConvGeom& g = p.cv;
  g.C = C; g.Fq = F2; g.TR = 128 / F2; g.Tq = T2; g.tiles_per_batch = (T2 + g.TR - 1) / g.TR; g.a_t_mul = 2; g.ntaps = 9;
  for (int t = 0; t < 9; ++t) { g.tap_t[t] = t / 3 - 1; g.tap_f[t] = t % 3 - 1; g.tap_k[t] = t * C; }
  g.To = T2; g.Fo = F2; g.ts = 1; g.fs = 1; g.pt = 0; g.pf = 0;
  p.M = B * g.tiles_per_batch; p.N = C; p.K = 9 * (C / BK); p.splits = 1; p.k_per_split = p.K;
  p.epi.out0 = out; p.epi.bias = bias; p.epi.lengths = len_out; p.epi.ldo = C; p.epi.rows_per_batch = T2 * F2;
  p.epi.rows_inner = F2; p.epi.alpha = 1.f;
  const bool wide = (C % 256 == 0);
  CUtensorMap ta, tb;
  int rc = conv_map4d(&ta, in, C, F1, T1, B, 64, F2, g.TR, 2);
  if (rc) return rc;
  uint64_t dims[2] = {(uint64_t)9 * C, (uint64_t)C}, strides[1] = {(uint64_t)9 * C * 2};
  uint32_t box[2] = {(uint32_t)BK, (uint32_t)((wide ? 256 : 128) / kCluster)}, es[2] = {1, 1};
  rc = make_tmap(&tb, w_bf16, 2, 2, dims, strides, box, es, 1);
  if (rc) return rc;
  if (wide) return launch_tc<256, 0, 0, B2C_EPI_RELU_MASK, 1>(ta, tb, p, stream);
  return launch_tc<128, 0, 0, B2C_EPI_RELU_MASK, 1>(ta, tb, p, stream);
}

// dx[b,t1,f1,ci] = relu'(x[b,t1,f1,ci]) * sum_{taps,co} dy[b,(t1+1-kh)/2,(f1+1-kw)/2,co] * W[co,ci,kh,kw]
// decomposed into the four (t1,f1) parity classes (1, 2, 2 and 4 taps): each a dense shifted-window implicit GEMM.
int conv3x3s2_bwd_data(const void* dy, const void* wd_bf16, const void* x, void* dx, int B, int T1, int F1, int C,
                       cudaStream_t stream) {
  const int T2 = (T1 - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  B2C_REQUIRE(C % 64 == 0 && F2 <= 128 && B > 0, "conv3x3s2_bwd_data: unsupported shape");
  const bool wide = (C % 256 == 0);
  CUtensorMap tb;
  {
    uint64_t dims[2] = {(uint64_t)9 * C, (uint64_t)C}, strides[1] = {(uint64_t)9 * C * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)((wide ? 256 : 128) / kCluster)}, es[2] = {1, 1};
    int rc = make_tmap(&tb, wd_bf16, 2, 2, dims, strides, box, es, 1);
    if (rc) return rc;
  }
  for (int pt = 0; pt < 2; ++pt) {
    for (int pf = 0; pf < 2; ++pf) {
      const int Tc = (T1 - pt + 1) / 2, Fc = (F1 - pf + 1) / 2;
      if (Tc <= 0 || Fc <= 0) continue;
      GemmParams p;
      memset(&p, 0, sizeof(p));
      ConvGeom& g = p.cv;
      g.C = C; g.Fq = Fc; g.TR = 128 / Fc; g.Tq = Tc; g.tiles_per_batch = (Tc + g.TR - 1) / g.TR; g.a_t_mul = 1;
      g.ntaps = 0;
      for (int kh = 0; kh < 3; ++kh) {
        if (((pt + 1 - kh) & 1) != 0) continue;
        for (int kw = 0; kw < 3; ++kw) {
          if (((pf + 1 - kw) & 1) != 0) continue;
          g.tap_t[g.ntaps] = (pt + 1 - kh) / 2;
          g.tap_f[g.ntaps] = (pf + 1 - kw) / 2;
          g.tap_k[g.ntaps] = (kh * 3 + kw) * C;
          ++g.ntaps;
        }
      }
      g.To = T1; g.Fo = F1; g.ts = 2; g.fs = 2; g.pt = pt; g.pf = pf;
      p.M = B * g.tiles_per_batch; p.N = C; p.K = g.ntaps * (C / BK); p.splits = 1; p.k_per_split = p.K;
      p.epi.out0 = dx; p.epi.aux = x; p.epi.ldo = C; p.epi.alpha = 1.f;
      CUtensorMap ta;
      int rc = conv_map4d(&ta, dy, C, F2, T2, B, 64, Fc, g.TR, 1);
      if (rc) return rc;
      rc = wide ? launch_tc<256, 0, 0, B2C_EPI_RELU_BWD, 2>(ta, tb, p, stream)
                : launch_tc<128, 0, 0, B2C_EPI_RELU_BWD, 2>(ta, tb, p, stream);
      if (rc) return rc;
    }
  }
  return B2C_OK;
}

// dw[co, tap*C + ci] += sum_{b,t2,f2} dy[b,t2,f2,co] * x[b,2t2+kh-1,2f2+kw-1,ci]
int conv3x3s2_bwd_weight(const void* dy, const void* x, float* dw, int B, int T1, int F1, int C, cudaStream_t stream) {
  const int T2 = (T1 - 1) / 2 + 1, F2 = (F1 - 1) / 2 + 1;
  B2C_REQUIRE(C % 128 == 0 && F2 <= 64 && B > 0, "conv3x3s2_bwd_weight: unsupported shape");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  ConvGeom& g = p.cv;
  g.C = C; g.Fq = F2; g.TR = 64 / F2; g.Tq = T2; g.tiles_per_batch = (T2 + g.TR - 1) / g.TR;
  p.M = C; p.N = C; p.K = B * g.tiles_per_batch;
  // 256-wide tiles halve the operand traffic per flop (this GEMM is bound by L2 -> SM bandwidth); the contraction is
  // split so that (pair-tiles x 9 taps x k-splits) still fills the CTA pairs
  const bool wide = (C % 256 == 0);
  const int bn = wide ? 256 : 128;
  const int pair_tiles = ((C / BM + kCluster - 1) / kCluster) * (C / bn);
  int ksplit = (num_sms() / kCluster) / (pair_tiles * 9);
  if (ksplit < 1) ksplit = 1;
  p.k_per_split = (p.K + ksplit - 1) / ksplit;
  ksplit = (p.K + p.k_per_split - 1) / p.k_per_split;
  p.splits = 9 * ksplit;
  p.epi.out0 = dw; p.epi.ldo = 9 * C; p.epi.alpha = 1.f;
  CUtensorMap ta, tb;
  int rc = conv_map4d(&ta, dy, C, F2, T2, B, 64, F2, g.TR, 1);
  if (rc) return rc;
  rc = conv_map4d(&tb, x, C, F1, T1, B, 64, F2, g.TR, 2);
  if (rc) return rc;
  if (wide) return launch_tc<256, 1, 1, B2C_EPI_ACCUM_F32, 3>(ta, tb, p, stream);
  return launch_tc<128, 1, 1, B2C_EPI_ACCUM_F32, 3>(ta, tb, p, stream);
}

}  // namespace b2c


extern "C" int b2c_gemm(const void* A, const void* B, int M, int N, int K, int lda, int ldb, int a_mn_major,
                        int b_mn_major, int epilogue_mode, const b2c_epilogue* epi, void* stream) {
  return b2c::gemm(A, B, M, N, K, lda, ldb, a_mn_major, b_mn_major, epilogue_mode, epi, (cudaStream_t)stream);
}

extern "C" int b2c_conv3x3s2_fwd(const void* in_bf16, const void* w_bf16, const float* bias, const int32_t* len_out,
                                 void* out_bf16, int B, int T1, int F1, int C, void* stream) {
  return b2c::conv3x3s2_fwd(in_bf16, w_bf16, bias, len_out, out_bf16, B, T1, F1, C, (cudaStream_t)stream);
}
extern "C" int b2c_conv3x3s2_bwd_data(const void* dy_bf16, const void* wd_bf16, const void* x_bf16, void* dx_bf16, int B,
                                      int T1, int F1, int C, void* stream) {
  return b2c::conv3x3s2_bwd_data(dy_bf16, wd_bf16, x_bf16, dx_bf16, B, T1, F1, C, (cudaStream_t)stream);
}
extern "C" int b2c_conv3x3s2_bwd_weight(const void* dy_bf16, const void* x_bf16, float* dw, int B, int T1, int F1, int C,
                                        void* stream) {
  return b2c::conv3x3s2_bwd_weight(dy_bf16, x_bf16, dw, B, T1, F1, C, (cudaStream_t)stream);
}
