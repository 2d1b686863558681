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

namespace atc {
constexpr int BQ = 128, BKV = 64, DK = 64, BAND = 192;
constexpr int kStages = 3;
constexpr int kKBytes = BKV * DK * 2, kVBytes = BKV * DK * 2, kPBytes = BAND * DK * 2;
constexpr int kStageBytes = kKBytes + kVBytes + kPBytes;  // 40 KB
constexpr int kQBytes = BQ * DK * 2;                      // 16 KB each (q+u, q+v)
constexpr int kProbBytes = BQ * BKV * 2;                  // 16 KB
constexpr int kOffQu = 0, kOffQv = kQBytes, kOffProb = 2 * kQBytes, kOffStage = 2 * kQBytes + kProbBytes;
constexpr int kOffBar = kOffStage + kStages * kStageBytes;
constexpr int kOffMax = kOffBar + 256;  // float [2][128] row-max / row-sum exchange between the two halves of a row
constexpr int kSmem = kOffMax + 2 * 128 * 4 + 1024;
constexpr int kThreads = 128 + 8 * 32;
constexpr uint32_t kTmemS = 0, kTmemR = 64, kTmemO = 256;  // column offsets
}  // namespace atc

// 64 columns of fp32 for 32 lanes
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

// write 8 bf16 (one 16-byte chunk `chunk` of row `row`) into a [rows x 64] K-major tile with the 128 B UMMA swizzle
__device__ __forceinline__ void st_sw128(uint8_t* tile, int row, int chunk, uint4 v) {
  *reinterpret_cast<uint4*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
}

__global__ void __launch_bounds__(atc::kThreads, 1)
relpos_attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                          const __grid_constant__ CUtensorMap tmP, const AttnArgs a, bf16* __restrict__ ctx,
                          float* __restrict__ ctx32, float* __restrict__ lse_out) {
  using namespace atc;
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
    mbar_init(p_ready, 8);
    mbar_init(o_full, 1);
    mbar_init(sr_free, 8);
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
                    idesc_o, k > 0 ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&empty_bar[stage]);  // K / P-band (consumed by SR_j) and V (by PV_j) of this stage are free
        stage = nstage;
        phase = nphase;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;              // TMEM lane quarter (= warp id % 4)
    const int half = (warp - 4) >> 2;    // which 32 key columns / 32 output columns of the row this thread owns
    const int il = q * 32 + lane;        // row inside the tile = TMEM lane
    const int row = i0 + il;
    const bool rvalid = row < len;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int cbase = 96 - 32 * q;       // warp-uniform part of the skew: R column = cbase + (31 - lane) + j_local
    const int sh = 31 - lane;
    const float sc2 = a.scale * kLog2e;
    const uint64_t rbase = attn_row_base(a, b, h, row);
    float* xch = reinterpret_cast<float*>(smem + kOffMax);  // [2][128]
    float o[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) o[k] = 0.f;
    float m_run = -1e30f, l_run = 0.f;
    uint8_t* prob = smem + kOffProb;

    for (int jt = 0; jt < n_kv; ++jt) {
      const int j0 = jt * BKV + half * 32;
      mbar_wait(sr_full, jt & 1);
      tc_fence_after();
      float s[32];
      {
        uint32_t sa[32], x[64];
        tmem_ld32(tlane + kTmemS + half * 32, sa);
        tmem_ld64(tlane + kTmemR + (uint32_t)(cbase + half * 32), x);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sr_free);
        // barrel shift left by sh = 31 - lane: afterwards x[k] = R[cbase + half*32 + sh + k]
#pragma unroll
        for (int k = 0; k < 48; ++k) x[k] = (sh & 16) ? x[k + 16] : x[k];
#pragma unroll
        for (int k = 0; k < 40; ++k) x[k] = (sh & 8) ? x[k + 8] : x[k];
#pragma unroll
        for (int k = 0; k < 36; ++k) x[k] = (sh & 4) ? x[k + 4] : x[k];
#pragma unroll
        for (int k = 0; k < 34; ++k) x[k] = (sh & 2) ? x[k + 2] : x[k];
#pragma unroll
        for (int k = 0; k < 32; ++k) x[k] = (sh & 1) ? x[k + 1] : x[k];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float v = (__uint_as_float(sa[k]) + __uint_as_float(x[k])) * sc2;
          s[k] = (rvalid && j0 + k < len) ? v : -1e30f;
        }
      }
      float mx = -1e30f;
#pragma unroll
      for (int k = 0; k < 32; ++k) mx = fmaxf(mx, s[k]);
      // row max across the two column halves: smem exchange + named barrier of the quarter's two warps
      xch[half * 128 + il] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      mx = fmaxf(mx, xch[(half ^ 1) * 128 + il]);
      const float m_new = fmaxf(m_run, mx);
      const float corr = exp2f(m_run - m_new);
      // O_{j-1} (computed against the previous running max) joins the register accumulator, then everything is rescaled
      if (jt > 0) {
        mbar_wait(o_full, (jt - 1) & 1);
        tc_fence_after();
        uint32_t ot[32];
        tmem_ld32(tlane + kTmemO + half * 32, ot);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) o[k] = (o[k] + __uint_as_float(ot[k])) * corr;
      }
      m_run = m_new;
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float p[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float sv = s[c * 8 + k];
          p[k] = (sv > -1e29f) ? exp2f(sv - m_new) : 0.f;
          rs += p[k];
        }
        if (a.drop_keep != 0) {
#pragma unroll
          for (int k = 0; k < 8; k += 2) {
            const uint32_t kp = attn_keep2(a, rbase, j0 + c * 8 + k);
            p[k] = (kp & 1u) ? p[k] * a.drop_scale : 0.f;
            p[k + 1] = (kp & 2u) ? p[k + 1] * a.drop_scale : 0.f;
          }
        }
        st_sw128(prob, il, half * 4 + c,
                 make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7])));
      }
      l_run = l_run * corr + rs;
      fence_proxy_async();
      tc_fence_before();
      // second barrier of the pair: both halves have read this tile's exchange slots before either overwrites them
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      if (lane == 0) mbar_arrive(p_ready);
    }
    if (n_kv > 0) {
      mbar_wait(o_full, (n_kv - 1) & 1);
      tc_fence_after();
      uint32_t ot[32];
      tmem_ld32(tlane + kTmemO + half * 32, ot);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) o[k] += __uint_as_float(ot[k]);
    }
    // total row sum = sum of the two halves (same running max)
    xch[half * 128 + il] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
    const float l_tot = l_run + xch[(half ^ 1) * 128 + il];
    if (row < T) {
      const float inv = (rvalid && l_tot > 0.f) ? 1.f / l_tot : 0.f;  // fully-masked query rows output exactly 0
      bf16* op = ctx + ((size_t)b * T + row) * a.d + h * DK + half * 32;
#pragma unroll
      for (int k = 0; k < 32; k += 8)
        *reinterpret_cast<uint4*>(op + k) =
            make_uint4(pack_bf16(o[k] * inv, o[k + 1] * inv), pack_bf16(o[k + 2] * inv, o[k + 3] * inv),
                       pack_bf16(o[k + 4] * inv, o[k + 5] * inv), pack_bf16(o[k + 6] * inv, o[k + 7] * inv));
      if (ctx32) {
        float* o32 = ctx32 + ((size_t)b * T + row) * a.d + h * DK + half * 32;
#pragma unroll
        for (int k = 0; k < 32; k += 4)
          *reinterpret_cast<float4*>(o32 + k) = make_float4(o[k] * inv, o[k + 1] * inv, o[k + 2] * inv, o[k + 3] * inv);
      }
      if (lse_out && half == 0) lse_out[((size_t)b * a.H + h) * T + row] = (rvalid && l_tot > 0.f) ? m_run + log2f(l_tot) : 0.f;
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
  static bool set = false;
  if (!set) {
    B2C_CUDA(cudaFuncSetAttribute(relpos_attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    set = true;
  }
  relpos_attn_fwd_tc_kernel<<<dim3((a.T + BQ - 1) / BQ, a.H, a.B), kThreads, kSmem, st>>>(tk, tv, tp, a, ctx, ctx32, lse);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

}  // namespace b2c
