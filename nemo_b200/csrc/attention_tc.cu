// K5 forward on 5th-gen tensor cores: Transformer-XL relative-position flash attention with tcgen05.mma,
// TMEM accumulators and TMA-staged K / V / positional-band tiles (head dim 64).
//
// One CTA = 128 query rows of one (batch, head); the key axis is walked in 64-key tiles:
//   warp 0      TMA producer: K [64 x 64], V [64 x 64] and the 192-row band of the projected positions p that the
//               128 x 64 score tile can touch (rel_shift folded into the band index, no [T, 2T-1] tensor anywhere)
//   warp 1      MMA issuer (one thread):  S = (q+u) K^T  (N = 64),  R = (q+v) P_band^T  (N = 192),  O_j = P_j V  (N = 64)
//   warp 2      TMEM allocator (512 columns: S | R | O)
//   warps 4..11 softmax: two threads per query row (TMEM lane), each owning 32 of the tile's 64 key columns and 32 of
//               the 64 output columns; row maxima are exchanged through smem + a 64-thread named barrier.
//               bd[i, j] = R[i, j + 127 - i]: the per-row skew is a warp-uniform TMEM column offset plus a 5-stage
//               register barrel shift by (31 - lane).
//               scores → mask(-1e4 semantics) → online softmax (exp2) → dropout → bf16 P tile written with the
//               128-byte UMMA swizzle into smem → PV on the tensor cores; O_j is pulled from TMEM and accumulated in
//               registers with the running-max correction.
// Reference: RelPositionMultiHeadAttention.forward (multi_head_attention.py:272-354) + forward_attention (:125-146).
#include "attn_common.cuh"

namespace b2c {

B2C_DEFINE_SEED_SETTER(set_seed_counter_attn_tc)


namespace atc {
constexpr int BQ = 128, BKV = 64, DK = 64, BAND = 192;
constexpr int kStages = 3;
constexpr int kKBytes = BKV * DK * 2, kVBytes = BKV * DK * 2, kPBytes = BAND * DK * 2;
constexpr int kStageBytes = kKBytes + kVBytes + kPBytes;  // 40 KB
constexpr int kQBytes = BQ * DK * 2;                      // 16 KB each (q+u, q+v)
constexpr int kProbBytes = BQ * BKV * 2;                  // 16 KB
constexpr int kOffQu = 0, kOffQv = kQBytes, kOffProb = 2 * kQBytes, kOffStage = 2 * kQBytes + kProbBytes;
constexpr int kOffBar = kOffStage + kStages * kStageBytes;
constexpr int kOffMax = kOffBar + 256;  // float [3][NT<=4][128] row-max / row-sum exchange between the threads of a row
constexpr int kSmem = kOffMax + 3 * 4 * 128 * 4 + 1024;
constexpr uint32_t kTmemS = 0, kTmemR = 64, kTmemO = 256;  // column offsets
}  // namespace atc

// 64 columns of fp32 for 32 lanes
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 2^x, single MUFU (exp2f() adds a denormal-range rescale: 3 more instructions per element)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
        "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
        "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
        "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
}

// N consecutive fp32 columns of this thread's TMEM lane, N in {16, 32, 48, 64}
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[N]) {
  static_assert(N == 16 || N == 32 || N == 48 || N == 64, "unsupported width");
  if constexpr (N == 16) {
    tmem_ld16(taddr, r);
  } else if constexpr (N == 32) {
    tmem_ld32(taddr, r);
  } else if constexpr (N == 64) {
    tmem_ld64(taddr, r);
  } else {
    tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
    tmem_ld16(taddr + 32, *reinterpret_cast<uint32_t(*)[16]>(&r[32]));
  }
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t (&r)[N]) {
  static_assert(N == 16 || N == 32, "unsupported width");
  if constexpr (N == 16) tmem_st16(taddr, r);
  else tmem_st32(taddr, r);
}
// Per-lane skew of the positional band: x holds CW + 32 consecutive band columns starting at the warp-uniform offset;
// afterwards x[k] = band[sh + k] for k < CW (5-stage barrel shift by sh in [0, 32))
template <int CW>
__device__ __forceinline__ void band_shift(uint32_t (&x)[CW + 32], int sh) {
  // (measured: the multiply-add form x[k]*m0 + x[k+s]*m1 on the FMA pipe is slower than these selects — the kernels are
  // bound by issue slots, not by the ALU pipe width)
#pragma unroll
  for (int k = 0; k < CW + 16; ++k) x[k] = (sh & 16) ? x[k + 16] : x[k];
#pragma unroll
  for (int k = 0; k < CW + 8; ++k) x[k] = (sh & 8) ? x[k + 8] : x[k];
#pragma unroll
  for (int k = 0; k < CW + 4; ++k) x[k] = (sh & 4) ? x[k + 4] : x[k];
#pragma unroll
  for (int k = 0; k < CW + 2; ++k) x[k] = (sh & 2) ? x[k + 2] : x[k];
#pragma unroll
  for (int k = 0; k < CW; ++k) x[k] = (sh & 1) ? x[k + 1] : x[k];
}

__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}

// write 8 bf16 (one 16-byte chunk `chunk` of row `row`) into a [rows x 64] K-major tile with the 128 B UMMA swizzle
__device__ __forceinline__ void st_sw128(uint8_t* tile, int row, int chunk, uint4 v) {
  sts128(smem_u32(tile) + row * 128 + ((chunk ^ (row & 7)) << 4), v);
}

// NT softmax threads per query row (NT * 4 softmax warps).  2 and 4 measure the same on B200: the kernel is bound by the
// instruction issue rate of the softmax / skew / dropout arithmetic, not by latency
// SKS (skew through shared memory): instead of the 5-stage register barrel shift (190 selects per thread and key tile), every
// softmax thread parks the 64 band columns it pulled from TMEM in its own padded shared-memory row (16 x STS.128) and reads
// the 32 it needs back at its lane-dependent offset (32 x LDS.32, bank-conflict free: row pitch 68 words).  The 68 KB this
// takes are paid for with one TMA stage (2 instead of 3).  Selected with B2C_ATTN_SKEW=smem (NT = 2 only).
template <int NT, bool SKS>
struct FwdLayout {
  static constexpr int kStg = SKS ? 2 : atc::kStages;
  static constexpr int kRowW = atc::BKV / NT + 32 + 4;  // words per softmax thread
  static constexpr int kBar = atc::kOffStage + kStg * atc::kStageBytes;
  static constexpr int kMax = kBar + 256;
  static constexpr int kSkew = kMax + 3 * 4 * 128 * 4;
  static constexpr int kTotal = kSkew + (SKS ? NT * 128 * kRowW * 4 : 0) + 1024;
};

template <int NT, bool SKS = false>
__global__ void __launch_bounds__(128 + NT * 128, 1)
relpos_attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                          const __grid_constant__ CUtensorMap tmP, const AttnArgs a, bf16* __restrict__ ctx,
                          float* __restrict__ ctx32, float* __restrict__ lse_out) {
  using namespace atc;
  using FL = FwdLayout<NT, SKS>;
  constexpr int kStages = FL::kStg;   // shadows atc::kStages
  constexpr int kOffBar = FL::kBar;   // ... and the offsets that depend on it
  constexpr int kOffMax = FL::kMax;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* sr_full = empty_bar + kStages;  // S and R of the current key tile are in TMEM
  uint64_t* p_ready = sr_full + 1;          // softmax has consumed O of the previous tile and published P
  uint64_t* o_full = p_ready + 1;           // O_j = P_j V is in TMEM
  uint64_t* sr_free = o_full + 1;           // S / R of the current tile are in registers: the next tile's may be issued
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sr_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int T = a.T, ld = 3 * a.d;
  const int len = a.lengths ? min(a.lengths[b], T) : T;
  const int n_kv = (len + BKV - 1) / BKV;  // keys >= len contribute exactly 0

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmP);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(sr_full, 1);
    mbar_init(p_ready, 4 * NT);
    mbar_init(o_full, 1);
    mbar_init(sr_free, 4 * NT);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  // (q + u), (q + v) → bf16 → swizzled K-major A tiles (all 256 threads; rows >= T are zero)
  {
    const bf16* qb = a.qkv + (size_t)b * T * ld + h * DK;
    const float* bu = a.bias_u + h * DK;
    const float* bv = a.bias_v + h * DK;
    for (int idx = threadIdx.x; idx < BQ * 8; idx += blockDim.x) {
      const int r = idx >> 3, ck = idx & 7;
      uint4 qu = make_uint4(0, 0, 0, 0), qv = make_uint4(0, 0, 0, 0);
      if (i0 + r < T) {
        const uint4 raw = *reinterpret_cast<const uint4*>(qb + (size_t)(i0 + r) * ld + ck * 8);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        uint32_t ou[4], ov[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16(w[k]);
          const int c = ck * 8 + 2 * k;
          ou[k] = pack_bf16(f.x + bu[c], f.y + bu[c + 1]);
          ov[k] = pack_bf16(f.x + bv[c], f.y + bv[c + 1]);
        }
        qu = make_uint4(ou[0], ou[1], ou[2], ou[3]);
        qv = make_uint4(ov[0], ov[1], ov[2], ov[3]);
      }
      st_sw128(smem + kOffQu, r, ck, qu);
      st_sw128(smem + kOffQv, r, ck, qv);
    }
  }
  fence_proxy_async();  // generic-proxy smem writes → visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int jt = 0; jt < n_kv; ++jt) {
        const int j0 = jt * BKV;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + kOffStage + stage * kStageBytes;
        mbar_expect_tx(&full_bar[stage], kStageBytes);
        tma_load_3d(st, &tmK, &full_bar[stage], a.d + h * DK, j0, b);
        tma_load_3d(st + kKBytes, &tmV, &full_bar[stage], 2 * a.d + h * DK, j0, b);
        tma_load_2d(st + kKBytes + kVBytes, &tmP, &full_bar[stage], h * DK, T - 1 - (i0 + BQ - 1) + j0);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_kv > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_r = umma_idesc_bf16(BQ, BAND, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(BQ, DK, 0, 1);  // B = V tile, MN-major (dcol contiguous)
      const uint32_t s_qu = smem_u32(smem + kOffQu), s_qv = smem_u32(smem + kOffQv), s_pr = smem_u32(smem + kOffProb);
      auto issue_sr = [&](int stage) {
        const uint32_t sk = smem_u32(smem + kOffStage + stage * kStageBytes);
        const uint32_t sp = sk + kKBytes + kVBytes;
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) {
          umma_bf16(tmem_base + kTmemS, umma_desc_sw128(s_qu + k * 32, 16, 1024), umma_desc_sw128(sk + k * 32, 16, 1024), idesc_s,
                    k > 0 ? 1u : 0u);
          umma_bf16(tmem_base + kTmemR, umma_desc_sw128(s_qv + k * 32, 16, 1024), umma_desc_sw128(sp + k * 32, 16, 1024), idesc_r,
                    k > 0 ? 1u : 0u);
        }
        umma_commit(sr_full);
      };
      int stage = 0, phase = 0;
      mbar_wait(&full_bar[0], 0);
      tc_fence_after();
      issue_sr(0);
      for (int jt = 0; jt < n_kv; ++jt) {
        int nstage = stage + 1, nphase = phase;
        if (nstage == kStages) { nstage = 0; nphase ^= 1; }
        if (jt + 1 < n_kv) {
          // as soon as the softmax warps hold S_j / R_j in registers, the next tile's scores go to the tensor cores:
          // they are computed while the softmax / exp / dropout of tile j is still running
          mbar_wait(sr_free, jt & 1);
          mbar_wait(&full_bar[nstage], nphase);
          tc_fence_after();
          issue_sr(nstage);
        }
        mbar_wait(p_ready, jt & 1);
        tc_fence_after();
        const uint32_t sv = smem_u32(smem + kOffStage + stage * kStageBytes) + kKBytes;
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          umma_bf16(tmem_base + kTmemO, umma_desc_sw128(s_pr + k * 32, 16, 1024), umma_desc_sw128(sv + k * 2048, 64 * BKV * 2, 1024),
                    idesc_o, (jt > 0 || k > 0) ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&empty_bar[stage]);  // K / P-band (consumed by SR_j) and V (by PV_j) of this stage are free
        stage = nstage;
        phase = nphase;
      }
    }
  } else if (warp >= 4) {
    constexpr int CW = BKV / NT;         // key columns (and output columns) of the row owned by this thread
    const int q = warp & 3;              // TMEM lane quarter (= warp id % 4)
    const int cq = (warp - 4) >> 2;      // which CW-column slice
    const int il = q * 32 + lane;        // row inside the tile = TMEM lane
    const int row = i0 + il;
    const bool rvalid = row < len;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int cbase = 96 - 32 * q;       // warp-uniform part of the skew: R column = cbase + (31 - lane) + j_local
    const int sh = 31 - lane;
    const float sc2 = a.scale * kLog2e;
    const uint64_t rbase = attn_row_base(a, b, h, row);
    float* xch = reinterpret_cast<float*>(smem + kOffMax);  // [3][NT][128]
    // Online softmax in the exp2 domain on the raw scores v = S + R: p = 2^(v*sc2 - m).  O accumulates in TMEM across
    // the key tiles; the reference maximum m is only moved (and O / l rescaled) when the running maximum has grown by
    // more than 2^8 since it was set, so most tiles touch neither O nor the o_full barrier.  The dropout scale 1/(1-p)
    // is applied once, with the final 1/l.
    float m_ref = -1e30f, l_run = 0.f, mx_prev = -1e30f;
    uint8_t* prob = smem + kOffProb;
    const uint32_t keep_hi = dropout_keep_hi(a.drop_keep);  // 16-bit lanes of the hash word compared in place

    for (int jt = 0; jt < n_kv; ++jt) {
      const int j0 = jt * BKV + cq * CW;
      mbar_wait(sr_full, jt & 1);
      tc_fence_after();
      float v[CW];
      {
        uint32_t sa[CW], x[CW + 32];
        tmem_ld_cols<CW>(tlane + kTmemS + cq * CW, sa);
        tmem_ld_cols<CW + 32>(tlane + kTmemR + (uint32_t)(cbase + cq * CW), x);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sr_free);
        if constexpr (SKS) {
          const uint32_t rowb = smem_u32(smem + FL::kSkew) + (uint32_t)((((warp - 4) * 32 + lane) * FL::kRowW) * 4);
#pragma unroll
          for (int c4 = 0; c4 < (CW + 32) / 4; ++c4)
            sts128(rowb + c4 * 16, make_uint4(x[4 * c4], x[4 * c4 + 1], x[4 * c4 + 2], x[4 * c4 + 3]));
#pragma unroll
          for (int k = 0; k < CW; ++k) x[k] = lds32(rowb + (uint32_t)((sh + k) * 4));
        } else {
          band_shift<CW>(x, sh);
        }
#pragma unroll
        for (int k = 0; k < CW; ++k) v[k] = __uint_as_float(sa[k]) + __uint_as_float(x[k]);
      }
      if (jt * BKV + BKV > len) {  // ragged last tile (CTA-uniform): keys beyond the length get probability 0
#pragma unroll
        for (int k = 0; k < CW; ++k) v[k] = (j0 + k < len) ? v[k] : -1e30f;
      }
      float mx = v[0];
#pragma unroll
      for (int k = 1; k < CW; ++k) mx = fmaxf(mx, v[k]);
      mx *= sc2;
      // Row maximum across the NT column slices.  Only tile 0 synchronises the warps of a quarter; afterwards the
      // reference follows the row maximum of the PREVIOUS tile (all slices read the same NT values, so they always
      // agree on it): 2^(v - m_ref) may then exceed 2^8 by however much this tile tops the last one — harmless in fp32 /
      // bf16, and it takes every named barrier out of the steady-state loop.
      float* slot = xch + (jt % 3) * (NT * 128);
      slot[cq * 128 + il] = mx;
      if (jt == 0) {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(NT * 32) : "memory");
        m_ref = mx;
#pragma unroll
        for (int o = 1; o < NT; ++o) m_ref = fmaxf(m_ref, slot[((cq + o) % NT) * 128 + il]);
      } else {
        // PV_{j-1} has retired: every softmax warp finished tile j-1 (its maximum is visible) and the P tile is free
        mbar_wait(o_full, (jt - 1) & 1);
        const float* pslot = xch + ((jt - 1) % 3) * (NT * 128);
        float seen = mx_prev;
#pragma unroll
        for (int o = 1; o < NT; ++o) seen = fmaxf(seen, pslot[((cq + o) % NT) * 128 + il]);
        if (__any_sync(0xffffffffu, seen > m_ref + 8.f)) {
          // move the reference maximum of the rows that need it: O (TMEM, complete up to tile j-1) and l are rescaled
          const float m_new = (seen > m_ref + 8.f) ? seen : m_ref;
          const float corr = ex2_approx(m_ref - m_new);
          m_ref = m_new;
          l_run *= corr;
          tc_fence_after();
          uint32_t ot[CW];
          tmem_ld_cols<CW>(tlane + kTmemO + cq * CW, ot);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < CW; ++k) ot[k] = __float_as_uint(__uint_as_float(ot[k]) * corr);
          tmem_st_cols<CW>(tlane + kTmemO + cq * CW, ot);
          tmem_st_wait();
        }
      }
      mx_prev = mx;
      float rs = 0.f;
      uint32_t kbase = 0;
#pragma unroll
      for (int c = 0; c < CW / 8; ++c) {
        float p[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          p[k] = ex2_approx(fmaf(v[c * 8 + k], sc2, -m_ref));
          rs += p[k];
        }
        if (a.drop_keep != 0) {
          if ((c & 1) == 0) kbase = attn_keep_base(a, rbase, j0 + c * 8);
#pragma unroll
          for (int k = 0; k < 8; k += 2) {
            uint32_t w0, w1;
            dropout_pair(kbase, (c & 1) * 4 + (k >> 1), w0, w1);
            p[k] = (w0 < keep_hi) ? p[k] : 0.f;
            p[k + 1] = (w1 < keep_hi) ? p[k + 1] : 0.f;
          }
        }
        st_sw128(prob, il, cq * (CW / 8) + c,
                 make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7])));
      }
      l_run += rs;
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    float o[CW];
#pragma unroll
    for (int k = 0; k < CW; ++k) o[k] = 0.f;
    if (n_kv > 0) {
      mbar_wait(o_full, (n_kv - 1) & 1);
      tc_fence_after();
      uint32_t ot[CW];
      tmem_ld_cols<CW>(tlane + kTmemO + cq * CW, ot);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < CW; ++k) o[k] = __uint_as_float(ot[k]);
    }
    // total row sum = sum of the slices (same reference maximum)
    xch[cq * 128 + il] = l_run;
    asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(NT * 32) : "memory");
    float l_tot = l_run;
#pragma unroll
    for (int o2 = 1; o2 < NT; ++o2) l_tot += xch[((cq + o2) % NT) * 128 + il];
    if (row < T) {
      const float inv = (rvalid && l_tot > 0.f) ? a.drop_scale / l_tot : 0.f;  // fully-masked query rows output exactly 0
      bf16* op = ctx + ((size_t)b * T + row) * a.d + h * DK + cq * CW;
#pragma unroll
      for (int k = 0; k < CW; k += 8)
        *reinterpret_cast<uint4*>(op + k) =
            make_uint4(pack_bf16(o[k] * inv, o[k + 1] * inv), pack_bf16(o[k + 2] * inv, o[k + 3] * inv),
                       pack_bf16(o[k + 4] * inv, o[k + 5] * inv), pack_bf16(o[k + 6] * inv, o[k + 7] * inv));
      if (ctx32) {
        float* o32 = ctx32 + ((size_t)b * T + row) * a.d + h * DK + cq * CW;
#pragma unroll
        for (int k = 0; k < CW; k += 4)
          *reinterpret_cast<float4*>(o32 + k) = make_float4(o[k] * inv, o[k + 1] * inv, o[k + 2] * inv, o[k + 3] * inv);
      }
      if (lse_out && cq == 0) lse_out[((size_t)b * a.H + h) * T + row] = (rvalid && l_tot > 0.f) ? m_ref + log2f(l_tot) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

int launch_relpos_attn_fwd_tc(const AttnArgs& a, bf16* ctx, float* ctx32, float* lse, cudaStream_t st) {
  using namespace atc;
  B2C_REQUIRE(a.d % 8 == 0 && a.d / a.H == DK, "relpos_attn_fwd_tc: head dim must be 64");
  CUtensorMap tk, tv, tp;
  {
    // q|k|v rows of one batch item: dims {3d, T, B}; a 64-key box per head slice; rows >= T are zero-filled
    uint64_t dims[3] = {(uint64_t)3 * a.d, (uint64_t)a.T, (uint64_t)a.B};
    uint64_t strides[2] = {(uint64_t)3 * a.d * 2, (uint64_t)a.T * 3 * a.d * 2};
    uint32_t box[3] = {(uint32_t)DK, (uint32_t)BKV, 1}, es[3] = {1, 1, 1};
    int rc = make_tmap(&tk, a.qkv, 2, 3, dims, strides, box, es, 1);
    if (rc) return rc;
    rc = make_tmap(&tv, a.qkv, 2, 3, dims, strides, box, es, 1);
    if (rc) return rc;
    uint64_t pd[2] = {(uint64_t)a.d, (uint64_t)(2 * a.T - 1)}, ps[1] = {(uint64_t)a.d * 2};
    uint32_t pb[2] = {(uint32_t)DK, (uint32_t)BAND}, pe[2] = {1, 1};
    rc = make_tmap(&tp, a.pos, 2, 2, pd, ps, pb, pe, 1);
    if (rc) return rc;
  }
  static int nt = 0;  // softmax threads per row: 2 (default) or 4 (B2C_ATTN_NT=4; measured equal: issue-slot bound)
  static int sks = 0;  // B2C_ATTN_SKEW=smem: positional-band skew through shared memory (NT = 2)
  if (nt == 0) {
    const char* e = getenv("B2C_ATTN_NT");
    nt = (e && atoi(e) == 4) ? 4 : 2;
    const char* k = getenv("B2C_ATTN_SKEW");
    sks = (nt == 2 && k && !strcmp(k, "smem")) ? 1 : 0;
    B2C_CUDA(cudaFuncSetAttribute(relpos_attn_fwd_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdLayout<2, false>::kTotal));
    B2C_CUDA(cudaFuncSetAttribute(relpos_attn_fwd_tc_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdLayout<4, false>::kTotal));
    B2C_CUDA(cudaFuncSetAttribute(relpos_attn_fwd_tc_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdLayout<2, true>::kTotal));
  }
  const dim3 grid((a.T + BQ - 1) / BQ, a.H, a.B);
  if (sks) relpos_attn_fwd_tc_kernel<2, true><<<grid, 128 + 2 * 128, FwdLayout<2, true>::kTotal, st>>>(tk, tv, tp, a, ctx, ctx32, lse);
  else if (nt == 2) relpos_attn_fwd_tc_kernel<2, false><<<grid, 128 + 2 * 128, FwdLayout<2, false>::kTotal, st>>>(tk, tv, tp, a, ctx, ctx32, lse);
  else relpos_attn_fwd_tc_kernel<4, false><<<grid, 128 + 4 * 128, FwdLayout<4, false>::kTotal, st>>>(tk, tv, tp, a, ctx, ctx32, lse);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}


// =============================================================================================== backward (query side)
// Same tiling as the forward: 128 query rows per CTA, 64-key tiles, 2 softmax threads per row.  Per tile the tensor cores
// produce S, R (positional band) and dP = dO V^T into TMEM; the softmax warps rebuild P = exp2(S - lse), form
// dS = P * (dP - delta) / sqrt(dk), publish it twice in shared memory — as the [128 x 64] A operand of dQ_u += dS K and,
// un-skewed (row i shifted by 127 - i), as the [128 x 192] A operand of dQ_v += dS~ P_band — and stage dS / P_dropped in
// HBM for the key-side and position-side kernels.  dQ_u and dQ_v accumulate in TMEM over the whole key loop.
namespace atb {
constexpr int BQ = 128, BKV = 64, DK = 64, BAND = 192;
constexpr int kStages = 2;
constexpr int kKBytes = BKV * DK * 2, kVBytes = BKV * DK * 2, kPBytes = BAND * DK * 2;
constexpr int kStageBytes = kKBytes + kVBytes + kPBytes;  // 40 KB
constexpr int kTile = BQ * DK * 2;                        // 16 KB: one [128 x 64] bf16 K-major tile
constexpr int kOffQu = 0, kOffQv = kTile, kOffDO = 2 * kTile, kOffDS = 3 * kTile, kOffDSU = 4 * kTile;  // dS~: 3 tiles
constexpr int kOffPD = 7 * kTile;  // P_dropped tile, only staged for the TMA store to the key-side workspace
constexpr int kOffStage = 8 * kTile;
constexpr int kOffBar = kOffStage + kStages * kStageBytes;
constexpr int kOffXch = kOffBar + 256;            // float [128] delta
constexpr int kOffXch2 = kOffXch + 128 * 4;       // uint4 [2][NT-1 <= 3][128]: boundary chunks of the un-skewed row slices
constexpr int kSmem = kOffXch2 + 2 * 3 * 128 * 16 + 1024;
constexpr uint32_t kTmS = 0, kTmR = 64, kTmDP = 256, kTmDQU = 320, kTmDQV = 384;
}  // namespace atb

template <int NT>
__global__ void __launch_bounds__(128 + NT * 128, 1)
relpos_attn_bwd_q_tc_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                            const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmDO,
                            const __grid_constant__ CUtensorMap tmWS, const __grid_constant__ CUtensorMap tmWP, const AttnArgs a,
                            const float* __restrict__ ctx32, const bf16* __restrict__ ctx, const bf16* __restrict__ dctx,
                            const float* __restrict__ lse, bf16* __restrict__ dqkv, float* __restrict__ dbias_u,
                            float* __restrict__ dbias_v, bf16* __restrict__ ws_ds, bf16* __restrict__ ws_p,
                            bf16* __restrict__ ws_qu, int Tp) {
  using namespace atb;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* srd_full = empty_bar + kStages;  // S, R, dP of the current tile are in TMEM
  uint64_t* srd_free = srd_full + 1;         // ... and have been pulled into registers
  uint64_t* ds_ready = srd_free + 1;         // dS / dS~ of the current tile are in shared memory
  uint64_t* dq_done = ds_ready + 1;          // the dQ MMAs that read them have retired
  uint64_t* do_full = dq_done + 1;           // dO tile landed
  uint64_t* st_done = do_full + 1;           // the TMA stores of the dS / P_dropped tiles have read shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(st_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int T = a.T, ld = 3 * a.d;
  const int len = a.lengths ? min(a.lengths[b], T) : T;
  const int n_kv = (len + BKV - 1) / BKV;
  const int n_kv_all = (T + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmP); tma_prefetch_desc(&tmDO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(srd_full, 1); mbar_init(srd_free, 4 * NT); mbar_init(ds_ready, 4 * NT); mbar_init(dq_done, 1); mbar_init(do_full, 1); mbar_init(st_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  {
    // (q+u), (q+v) A tiles; zero the un-skew tile (each row only ever writes its own fixed 64-column window)
    const bf16* qb = a.qkv + (size_t)b * T * ld + h * DK;
    const float* bu = a.bias_u + h * DK;
    const float* bv = a.bias_v + h * DK;
    for (int idx = threadIdx.x; idx < BQ * 8; idx += blockDim.x) {
      const int r = idx >> 3, ck = idx & 7;
      uint4 qu = make_uint4(0, 0, 0, 0), qv = make_uint4(0, 0, 0, 0);
      if (i0 + r < T) {
        const uint4 raw = *reinterpret_cast<const uint4*>(qb + (size_t)(i0 + r) * ld + ck * 8);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
        uint32_t ou[4], ov[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16(w[k]);
          const int c = ck * 8 + 2 * k;
          ou[k] = pack_bf16(f.x + bu[c], f.y + bu[c + 1]);
          ov[k] = pack_bf16(f.x + bv[c], f.y + bv[c + 1]);
        }
        qu = make_uint4(ou[0], ou[1], ou[2], ou[3]);
        qv = make_uint4(ov[0], ov[1], ov[2], ov[3]);
        if (ws_qu) *reinterpret_cast<uint4*>(ws_qu + ((size_t)b * T + i0 + r) * a.d + h * DK + ck * 8) = qu;
      }
      st_sw128(smem + kOffQu, r, ck, qu);
      st_sw128(smem + kOffQv, r, ck, qv);
    }
    for (int i = threadIdx.x; i < 3 * kTile / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(smem + kOffDSU)[i] = make_uint4(0, 0, 0, 0);
    // delta_i = sum_d dO[i,d] * O[i,d] (fp32 O when available): 16 threads per row, 4 columns each, shuffle-reduced
    float* dl = reinterpret_cast<float*>(smem + kOffXch);
    for (int idx = threadIdx.x; idx < BQ * 16; idx += blockDim.x) {  // blockDim % 16 == 0: a row never straddles warps
      const int r = idx >> 4, c4 = (idx & 15) * 4;
      float acc = 0.f;
      if (i0 + r < T) {
        const size_t off = ((size_t)b * T + i0 + r) * a.d + h * DK + c4;
        const uint2 dy = *reinterpret_cast<const uint2*>(dctx + off);
        const float2 y0 = unpack_bf16(dy.x), y1 = unpack_bf16(dy.y);
        if (ctx32) {
          const float4 o = *reinterpret_cast<const float4*>(ctx32 + off);
          acc = o.x * y0.x + o.y * y0.y + o.z * y1.x + o.w * y1.y;
        } else {
          const uint2 ox = *reinterpret_cast<const uint2*>(ctx + off);
          const float2 o0 = unpack_bf16(ox.x), o1 = unpack_bf16(ox.y);
          acc = o0.x * y0.x + o0.y * y0.y + o1.x * y1.x + o1.y * y1.y;
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if ((idx & 15) == 0) dl[r] = acc;
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && n_kv > 0) {
      mbar_expect_tx(do_full, kTile);
      tma_load_3d(smem + kOffDO, &tmDO, do_full, h * DK, i0, b);
      int stage = 0, phase = 0;
      for (int jt = 0; jt < n_kv; ++jt) {
        const int j0 = jt * BKV;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + kOffStage + stage * kStageBytes;
        mbar_expect_tx(&full_bar[stage], kStageBytes);
        tma_load_3d(st, &tmK, &full_bar[stage], a.d + h * DK, j0, b);
        tma_load_3d(st + kKBytes, &tmV, &full_bar[stage], 2 * a.d + h * DK, j0, b);
        tma_load_2d(st + kKBytes + kVBytes, &tmP, &full_bar[stage], h * DK, T - 1 - (i0 + BQ - 1) + j0);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_kv > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_r = umma_idesc_bf16(BQ, BAND, 0, 0);
      constexpr uint32_t idesc_q = umma_idesc_bf16(BQ, DK, 0, 1);  // B = K tile / P band as [k rows][dcol contiguous]
      const uint32_t s_qu = smem_u32(smem + kOffQu), s_qv = smem_u32(smem + kOffQv), s_do = smem_u32(smem + kOffDO);
      const uint32_t s_ds = smem_u32(smem + kOffDS), s_dsu = smem_u32(smem + kOffDSU);
      auto issue_srd = [&](int stage) {
        const uint32_t sk = smem_u32(smem + kOffStage + stage * kStageBytes);
        const uint32_t sv = sk + kKBytes, sp = sv + kVBytes;
#pragma unroll
        for (int k = 0; k < DK / 16; ++k) {
          umma_bf16(tmem_base + kTmS, umma_desc_sw128(s_qu + k * 32, 16, 1024), umma_desc_sw128(sk + k * 32, 16, 1024), idesc_s, k > 0);
          umma_bf16(tmem_base + kTmR, umma_desc_sw128(s_qv + k * 32, 16, 1024), umma_desc_sw128(sp + k * 32, 16, 1024), idesc_r, k > 0);
          umma_bf16(tmem_base + kTmDP, umma_desc_sw128(s_do + k * 32, 16, 1024), umma_desc_sw128(sv + k * 32, 16, 1024), idesc_s, k > 0);
        }
        umma_commit(srd_full);
      };
      int stage = 0, phase = 0;
      mbar_wait(do_full, 0);
      mbar_wait(&full_bar[0], 0);
      tc_fence_after();
      issue_srd(0);
      for (int jt = 0; jt < n_kv; ++jt) {
        int nstage = stage + 1, nphase = phase;
        if (nstage == kStages) { nstage = 0; nphase ^= 1; }
        if (jt + 1 < n_kv) {
          mbar_wait(srd_free, jt & 1);
          mbar_wait(&full_bar[nstage], nphase);
          tc_fence_after();
          issue_srd(nstage);
        }
        mbar_wait(ds_ready, jt & 1);
        tc_fence_after();
        const uint32_t sk = smem_u32(smem + kOffStage + stage * kStageBytes);
        const uint32_t sp = sk + kKBytes + kVBytes;
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)   // dQ_u += dS (128 x 64 keys) * K (64 keys x 64)
          umma_bf16(tmem_base + kTmDQU, umma_desc_sw128(s_ds + k * 32, 16, 1024), umma_desc_sw128(sk + k * 2048, 64 * BKV * 2, 1024),
                    idesc_q, (jt > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < BAND / 16; ++k)  // dQ_v += dS~ (128 x 192 band) * P_band (192 x 64)
          umma_bf16(tmem_base + kTmDQV, umma_desc_sw128(s_dsu + (k >> 2) * kTile + (k & 3) * 32, 16, 1024),
                    umma_desc_sw128(sp + k * 2048, 64 * BAND * 2, 1024), idesc_q, (jt > 0 || k > 0) ? 1u : 0u);
        umma_commit(dq_done);
        umma_commit(&empty_bar[stage]);
        stage = nstage;
        phase = nphase;
      }
    }
  } else if (warp == 3) {
    // staging for the key- and position-side kernels: the finished dS / P_dropped tiles leave as two TMA stores
    // (the tiles already have the 128-byte-swizzle layout; rows >= T and columns >= Tp are clipped by the tensor map)
    if (lane == 0) {
      const int bh = b * a.H + h;
      for (int jt = 0; jt < n_kv; ++jt) {
        mbar_wait(ds_ready, jt & 1);
        tma_store_3d(&tmWS, smem + kOffDS, jt * BKV, i0, bh);
        tma_store_3d(&tmWP, smem + kOffPD, jt * BKV, i0, bh);
        bulk_commit();
        bulk_wait_read<0>();
        mbar_arrive(st_done);
      }
      bulk_wait_all();
    }
  } else if (warp >= 4) {
    constexpr int CW = BKV / NT;   // key columns of the row owned by this thread (and dq columns in the epilogue)
    constexpr int PW = CW / 2;     // ... as packed bf16 pairs
    constexpr int NC = PW / 4 + 1; // 16-byte chunks the un-skewed slice can touch
    const int q = warp & 3, cq = (warp - 4) >> 2;
    const int il = q * 32 + lane, row = i0 + il;
    const bool rvalid = row < len;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int cbase = 96 - 32 * q, sh = 31 - lane;
    const float sc2 = a.scale * kLog2e;
    const uint64_t rbase = attn_row_base(a, b, h, row);
    float* xch = reinterpret_cast<float*>(smem + kOffXch);
    // delta_i = sum_d dO[i,d] * O[i,d] was put into xch[] by the whole CTA before the role split (coalesced loads)
    const float delta = xch[il];
    float lse2 = 1e30f;  // padded query rows: p = 0
    if (rvalid) lse2 = lse[((size_t)b * a.H + h) * T + row];
    const uint32_t keep_hi = dropout_keep_hi(a.drop_keep);  // 16-bit lanes of the hash word compared in place
    bf16* wds = ws_ds + ((size_t)(b * a.H + h) * T) * Tp;
    bf16* wp = ws_p + ((size_t)(b * a.H + h) * T) * Tp;
    uint8_t* s_ds = smem + kOffDS;
    uint8_t* s_dsu = smem + kOffDSU;

    for (int jt = 0; jt < n_kv; ++jt) {
      const int j0 = jt * BKV + cq * CW;
      mbar_wait(srd_full, jt & 1);
      tc_fence_after();
      float ds[CW], pd[CW];
      {
        uint32_t sa[CW], x[CW + 32], dpv[CW];
        tmem_ld_cols<CW>(tlane + kTmS + cq * CW, sa);
        tmem_ld_cols<CW + 32>(tlane + kTmR + (uint32_t)(cbase + cq * CW), x);
        tmem_ld_cols<CW>(tlane + kTmDP + cq * CW, dpv);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(srd_free);
        band_shift<CW>(x, sh);
        // p = 2^((S+R)*sc2 - lse) (lse = +huge for padded query rows → 0; padded keys of the ragged last tile → 0),
        // pd = dropout(p), ds = p * (dropout(dP) - delta) / sqrt(dk)
        const bool ragged = jt * BKV + BKV > len;  // CTA-uniform
        uint32_t kbase = 0;
#pragma unroll
        for (int k = 0; k < CW; k += 2) {
          float km0 = a.drop_scale, km1 = a.drop_scale;
          if (a.drop_keep != 0) {
            if ((k & 15) == 0) kbase = attn_keep_base(a, rbase, j0 + k);
            uint32_t w0, w1;
            dropout_pair(kbase, (k & 15) >> 1, w0, w1);
            km0 = (w0 < keep_hi) ? a.drop_scale : 0.f;
            km1 = (w1 < keep_hi) ? a.drop_scale : 0.f;
          }
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float v = __uint_as_float(sa[k + e]) + __uint_as_float(x[k + e]);
            if (ragged) v = (j0 + k + e < len) ? v : -1e30f;
            const float p = ex2_approx(fmaf(v, sc2, -lse2));
            const float km = e ? km1 : km0;
            pd[k + e] = p * km;
            ds[k + e] = (p * a.scale) * fmaf(__uint_as_float(dpv[k + e]), km, -delta);
          }
        }
      }
      // the previous tile's dQ MMAs and TMA stores must have finished reading the dS / dS~ / P tiles before they are
      // overwritten
      if (jt > 0) {
        mbar_wait(dq_done, (jt - 1) & 1);
        mbar_wait(st_done, (jt - 1) & 1);
      }
      // dS → A operand tile (K-major, swizzled); dS and P_dropped also go to HBM from these tiles (store warp)
      uint32_t dsp[PW];
#pragma unroll
      for (int k = 0; k < PW; ++k) dsp[k] = pack_bf16(ds[2 * k], ds[2 * k + 1]);
#pragma unroll
      for (int c = 0; c < CW / 8; ++c) {
        st_sw128(s_ds, il, cq * (CW / 8) + c, make_uint4(dsp[4 * c], dsp[4 * c + 1], dsp[4 * c + 2], dsp[4 * c + 3]));
        st_sw128(smem + kOffPD, il, cq * (CW / 8) + c,
                 make_uint4(pack_bf16(pd[8 * c], pd[8 * c + 1]), pack_bf16(pd[8 * c + 2], pd[8 * c + 3]),
                            pack_bf16(pd[8 * c + 4], pd[8 * c + 5]), pack_bf16(pd[8 * c + 6], pd[8 * c + 7])));
      }
      // un-skewed copy: dS~[il][x] with x = j_local + 127 - il (a fixed 64-column window per row; the rest stays zero).
      // x = cq*CW + cbase + sh + k: the 16-byte chunk part of the shift is address arithmetic, the remaining
      // (sh & 7) elements are a half-word funnel shift plus a 2-stage word barrel shift of the packed slice.  Adjacent
      // slices of a row meet inside one chunk: each thread hands its last (partial) chunk to the next one, which merges it.
      {
        const int es = sh & 7;
        const uint32_t odd16 = (uint32_t)(es & 1) * 16u, wsh = (uint32_t)es >> 1;
        uint32_t e[4 * NC + 1];
        e[0] = __funnelshift_l(0u, dsp[0], odd16);
#pragma unroll
        for (int k = 1; k < PW; ++k) e[k] = __funnelshift_l(dsp[k - 1], dsp[k], odd16);
        e[PW] = __funnelshift_l(dsp[PW - 1], 0u, odd16);
#pragma unroll
        for (int k = PW + 1; k <= 4 * NC; ++k) e[k] = 0u;
#pragma unroll
        for (int k = 4 * NC - 1; k >= 2; --k) e[k] = (wsh & 2u) ? e[k - 2] : e[k];
        if (wsh & 2u) e[0] = e[1] = 0u;
#pragma unroll
        for (int k = 4 * NC - 1; k >= 1; --k) e[k] = (wsh & 1u) ? e[k - 1] : e[k];
        if (wsh & 1u) e[0] = 0u;
        const int cid0 = cq * (CW / 8) + (cbase >> 3) + (sh >> 3);
        const uint32_t xbase = smem_u32(smem + kOffXch2) + (uint32_t)((jt & 1) * (NT - 1) * 128 * 16);
        if (cq < NT - 1)
          sts128(xbase + (uint32_t)((cq * 128 + il) * 16), make_uint4(e[4 * NC - 4], e[4 * NC - 3], e[4 * NC - 2], e[4 * NC - 1]));
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(NT * 32) : "memory");
        if (cq > 0) {
          const uint4 lo = lds128(xbase + (uint32_t)(((cq - 1) * 128 + il) * 16));
          e[0] |= lo.x; e[1] |= lo.y; e[2] |= lo.z; e[3] |= lo.w;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          if (c < NC - 1 || cq == NT - 1) {
            const int cid = cid0 + c;
            st_sw128(s_dsu + (cid >> 3) * kTile, il, cid & 7, make_uint4(e[4 * c], e[4 * c + 1], e[4 * c + 2], e[4 * c + 3]));
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_ready);
    }
    // key tiles entirely beyond the valid length: dS = P = 0 in the staging buffers
    if (row < T) {
      for (int j = n_kv * BKV + cq * CW; j < n_kv_all * BKV; j += BKV)
#pragma unroll
        for (int c = 0; c < CW / 8; ++c)
          if (j + c * 8 < Tp) {
            *reinterpret_cast<uint4*>(wds + (size_t)row * Tp + j + c * 8) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(wp + (size_t)row * Tp + j + c * 8) = make_uint4(0, 0, 0, 0);
          }
    }
    // dq = dQ_u + dQ_v; d(pos_bias_u/v) = column sums over rows
    float du[CW], dv[CW];
#pragma unroll
    for (int k = 0; k < CW; ++k) du[k] = dv[k] = 0.f;
    if (n_kv > 0) {
      mbar_wait(dq_done, (n_kv - 1) & 1);
      tc_fence_after();
      uint32_t t0[CW], t1[CW];
      tmem_ld_cols<CW>(tlane + kTmDQU + cq * CW, t0);
      tmem_ld_cols<CW>(tlane + kTmDQV + cq * CW, t1);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < CW; ++k) { du[k] = __uint_as_float(t0[k]); dv[k] = __uint_as_float(t1[k]); }
    }
    if (row < T) {
      bf16* op = dqkv + ((size_t)b * T + row) * ld + h * DK + cq * CW;
#pragma unroll
      for (int k = 0; k < CW; k += 8)
        *reinterpret_cast<uint4*>(op + k) =
            make_uint4(pack_bf16(du[k] + dv[k], du[k + 1] + dv[k + 1]), pack_bf16(du[k + 2] + dv[k + 2], du[k + 3] + dv[k + 3]),
                       pack_bf16(du[k + 4] + dv[k + 4], du[k + 5] + dv[k + 5]), pack_bf16(du[k + 6] + dv[k + 6], du[k + 7] + dv[k + 7]));
    }
#pragma unroll
    for (int k = 0; k < CW; ++k) {
      const float su = warp_sum(du[k]), sv2 = warp_sum(dv[k]);  // rows >= len contribute exact zeros
      if (lane == 0) {
        atomicAdd(dbias_u + h * DK + cq * CW + k, su);
        atomicAdd(dbias_v + h * DK + cq * CW + k, sv2);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

int launch_relpos_attn_bwd_q_tc(const AttnArgs& a, const bf16* ctx, const float* ctx32, const bf16* dctx, const float* lse,
                                bf16* dqkv, float* dbu, float* dbv, bf16* ws_ds, bf16* ws_p, bf16* ws_qu, int Tp, cudaStream_t st) {
  using namespace atb;
  B2C_REQUIRE(a.d % 8 == 0 && a.d / a.H == DK && Tp % 8 == 0, "relpos_attn_bwd_q_tc: head dim must be 64");
  CUtensorMap tk, tv, tp, tdo, tws, twp;
  {
    uint64_t dims[3] = {(uint64_t)3 * a.d, (uint64_t)a.T, (uint64_t)a.B};
    uint64_t strides[2] = {(uint64_t)3 * a.d * 2, (uint64_t)a.T * 3 * a.d * 2};
    uint32_t box[3] = {(uint32_t)DK, (uint32_t)BKV, 1}, es[3] = {1, 1, 1};
    int rc = make_tmap(&tk, a.qkv, 2, 3, dims, strides, box, es, 1);
    if (rc) return rc;
    rc = make_tmap(&tv, a.qkv, 2, 3, dims, strides, box, es, 1);
    if (rc) return rc;
    uint64_t pd[2] = {(uint64_t)a.d, (uint64_t)(2 * a.T - 1)}, ps[1] = {(uint64_t)a.d * 2};
    uint32_t pb[2] = {(uint32_t)DK, (uint32_t)BAND}, pe[2] = {1, 1};
    rc = make_tmap(&tp, a.pos, 2, 2, pd, ps, pb, pe, 1);
    if (rc) return rc;
    uint64_t dd[3] = {(uint64_t)a.d, (uint64_t)a.T, (uint64_t)a.B};
    uint64_t dstr[2] = {(uint64_t)a.d * 2, (uint64_t)a.T * a.d * 2};
    uint32_t dbox[3] = {(uint32_t)DK, (uint32_t)BQ, 1};
    rc = make_tmap(&tdo, dctx, 2, 3, dd, dstr, dbox, es, 1);
    if (rc) return rc;
    uint64_t wd[3] = {(uint64_t)Tp, (uint64_t)a.T, (uint64_t)a.B * a.H};
    uint64_t wstr[2] = {(uint64_t)Tp * 2, (uint64_t)a.T * Tp * 2};
    uint32_t wbox[3] = {(uint32_t)BKV, (uint32_t)BQ, 1};
    rc = make_tmap(&tws, ws_ds, 2, 3, wd, wstr, wbox, es, 1);
    if (rc) return rc;
    rc = make_tmap(&twp, ws_p, 2, 3, wd, wstr, wbox, es, 1);
    if (rc) return rc;
  }
  static int nt = 0;  // softmax threads per row: 2 (default) or 4 (B2C_ATTN_NT=4; measured equal: issue-slot bound)
  if (nt == 0) {
    const char* e = getenv("B2C_ATTN_NT");
    nt = (e && atoi(e) == 4) ? 4 : 2;
    B2C_CUDA(cudaFuncSetAttribute(relpos_attn_bwd_q_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    B2C_CUDA(cudaFuncSetAttribute(relpos_attn_bwd_q_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  }
  const dim3 grid((a.T + BQ - 1) / BQ, a.H, a.B);
  if (nt == 2)
    relpos_attn_bwd_q_tc_kernel<2><<<grid, 128 + 2 * 128, kSmem, st>>>(tk, tv, tp, tdo, tws, twp, a, ctx32, ctx, dctx, lse, dqkv, dbu, dbv,
                                                                       ws_ds, ws_p, ws_qu, Tp);
  else
    relpos_attn_bwd_q_tc_kernel<4><<<grid, 128 + 4 * 128, kSmem, st>>>(tk, tv, tp, tdo, tws, twp, a, ctx32, ctx, dctx, lse, dqkv, dbu, dbv,
                                                                       ws_ds, ws_p, ws_qu, Tp);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}


// =============================================================================================== backward (key side)
// dK[j,:] = sum_i dS[i,j] (q_i+u),  dV[j,:] = sum_i Pd[i,j] dO[i,:] from the staged bf16 dS / P_dropped tiles: a pure
// TMA → tcgen05 stream.  CTA = 128 keys of one (b,h); per 64-query tile both operands are used exactly as they lie in HBM
// (A = dS^T / Pd^T: keys contiguous → MN-major, B = (q+u) / dO: channels contiguous → MN-major).  2 CTAs per SM.
namespace akv {
constexpr int BK = 128, BQ = 64, DK = 64;
constexpr int kStages = 2;
constexpr int kHalf = BQ * 64 * 2;                    // 8 KB: [64 queries][64 keys or channels]
constexpr int kStageBytes = 6 * kHalf;                // dS (2 key chunks) | P (2) | q+u | dO = 48 KB
constexpr int kOffBar = kStages * kStageBytes;
constexpr int kSmem = kOffBar + 128 + 1024;
constexpr int kThreads = 256;
}  // namespace akv

__global__ void __launch_bounds__(akv::kThreads, 2)
relpos_attn_bwd_kv_tc_kernel(const __grid_constant__ CUtensorMap tmDS, const __grid_constant__ CUtensorMap tmPD,
                             const __grid_constant__ CUtensorMap tmQU, const __grid_constant__ CUtensorMap tmDO, const AttnArgs a,
                             bf16* __restrict__ dqkv) {
  using namespace akv;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_full = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * BK, h = blockIdx.y, b = blockIdx.z;
  const int T = a.T, ld = 3 * a.d;
  const int len = a.lengths ? min(a.lengths[b], T) : T;
  const int n_q = (j0 < len) ? (len + BQ - 1) / BQ : 0;  // keys beyond the length have dS = P = 0 everywhere

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDS); tma_prefetch_desc(&tmPD); tma_prefetch_desc(&tmQU); tma_prefetch_desc(&tmDO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, phase = 0;
      const int bh = b * a.H + h;
      for (int it = 0; it < n_q; ++it) {
        const int i0 = it * BQ;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * kStageBytes;
        mbar_expect_tx(&full_bar[stage], kStageBytes);
        tma_load_3d(st, &tmDS, &full_bar[stage], j0, i0, bh);
        tma_load_3d(st + kHalf, &tmDS, &full_bar[stage], j0 + 64, i0, bh);
        tma_load_3d(st + 2 * kHalf, &tmPD, &full_bar[stage], j0, i0, bh);
        tma_load_3d(st + 3 * kHalf, &tmPD, &full_bar[stage], j0 + 64, i0, bh);
        tma_load_3d(st + 4 * kHalf, &tmQU, &full_bar[stage], h * DK, i0, b);
        tma_load_3d(st + 5 * kHalf, &tmDO, &full_bar[stage], h * DK, i0, b);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_q > 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BK, DK, 1, 1);
      int stage = 0, phase = 0;
      for (int it = 0; it < n_q; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t s0 = smem_u32(smem + stage * kStageBytes);
#pragma unroll
        for (int k = 0; k < BQ / 16; ++k) {
          umma_bf16(tmem_base, umma_desc_sw128(s0 + k * 2048, kHalf, 1024), umma_desc_sw128(s0 + 4 * kHalf + k * 2048, kHalf, 1024),
                    idesc, (it > 0 || k > 0) ? 1u : 0u);
          umma_bf16(tmem_base + 64, umma_desc_sw128(s0 + 2 * kHalf + k * 2048, kHalf, 1024),
                    umma_desc_sw128(s0 + 5 * kHalf + k * 2048, kHalf, 1024), idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int j = j0 + q * 32 + lane;
    uint32_t dk[64], dv[32];
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    if (n_q > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      tmem_ld64(tlane, dk);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int k = 0; k < 64; ++k) dk[k] = 0u;
    }
    if (j < T) {
      bf16* op = dqkv + ((size_t)b * T + j) * ld + a.d + h * DK;
#pragma unroll
      for (int k = 0; k < 64; k += 8)
        *reinterpret_cast<uint4*>(op + k) = make_uint4(
            pack_bf16(__uint_as_float(dk[k]), __uint_as_float(dk[k + 1])), pack_bf16(__uint_as_float(dk[k + 2]), __uint_as_float(dk[k + 3])),
            pack_bf16(__uint_as_float(dk[k + 4]), __uint_as_float(dk[k + 5])), pack_bf16(__uint_as_float(dk[k + 6]), __uint_as_float(dk[k + 7])));
    }
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      if (n_q > 0) {
        tmem_ld32(tlane + 64 + hf * 32, dv);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) dv[k] = 0u;
      }
      if (j < T) {
        bf16* op = dqkv + ((size_t)b * T + j) * ld + 2 * a.d + h * DK + hf * 32;
#pragma unroll
        for (int k = 0; k < 32; k += 8)
          *reinterpret_cast<uint4*>(op + k) = make_uint4(
              pack_bf16(__uint_as_float(dv[k]), __uint_as_float(dv[k + 1])), pack_bf16(__uint_as_float(dv[k + 2]), __uint_as_float(dv[k + 3])),
              pack_bf16(__uint_as_float(dv[k + 4]), __uint_as_float(dv[k + 5])), pack_bf16(__uint_as_float(dv[k + 6]), __uint_as_float(dv[k + 7])));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

int launch_relpos_attn_bwd_kv_tc(const AttnArgs& a, const bf16* dctx, const bf16* ws_ds, const bf16* ws_p, const bf16* ws_qu,
                                 bf16* dqkv, int Tp, cudaStream_t st) {
  using namespace akv;
  CUtensorMap tds, tpd, tqu, tdo;
  {
    uint64_t dims[3] = {(uint64_t)Tp, (uint64_t)a.T, (uint64_t)a.B * a.H};
    uint64_t strides[2] = {(uint64_t)Tp * 2, (uint64_t)a.T * Tp * 2};
    uint32_t box[3] = {64, (uint32_t)BQ, 1}, es[3] = {1, 1, 1};
    int rc = make_tmap(&tds, ws_ds, 2, 3, dims, strides, box, es, 1);
    if (rc) return rc;
    rc = make_tmap(&tpd, ws_p, 2, 3, dims, strides, box, es, 1);
    if (rc) return rc;
    uint64_t dd[3] = {(uint64_t)a.d, (uint64_t)a.T, (uint64_t)a.B};
    uint64_t dstr[2] = {(uint64_t)a.d * 2, (uint64_t)a.T * a.d * 2};
    uint32_t dbox[3] = {(uint32_t)DK, (uint32_t)BQ, 1};
    rc = make_tmap(&tqu, ws_qu, 2, 3, dd, dstr, dbox, es, 1);
    if (rc) return rc;
    rc = make_tmap(&tdo, dctx, 2, 3, dd, dstr, dbox, es, 1);
    if (rc) return rc;
  }
  static bool set = false;
  if (!set) {
    B2C_CUDA(cudaFuncSetAttribute(relpos_attn_bwd_kv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    set = true;
  }
  relpos_attn_bwd_kv_tc_kernel<<<dim3((a.T + BK - 1) / BK, a.H, a.B), kThreads, kSmem, st>>>(tds, tpd, tqu, tdo, a, dqkv);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

}  // namespace b2c
