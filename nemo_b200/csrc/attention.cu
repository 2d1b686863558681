// K5: Transformer-XL relative-position multi-head attention, flash-style (nothing [T,T]-shaped is kept in the
// forward; the backward stages dS / P tiles once in bf16 for the key-side and position-side reductions).
//
//   S[i,j] = ((q_i + u) . k_j  +  (q_i + v) . p[T-1-i+j]) / sqrt(dk)          rel_shift folded into the band index
//   S = masked_fill(S, ~(valid_i & valid_j), -1e4); A = softmax(S); A = masked_fill(A, mask, 0); A = dropout(A)
//   ctx = A V
// Reference: RelPositionMultiHeadAttention.forward / rel_shift / forward_attention,
//            nemo/collections/asr/parts/submodules/multi_head_attention.py:125-146,259-354.
//
// Round-1 implementation: warp-level mma.sync.m16n8k16 (bf16, fp32 accumulate) with the positional band of
// 16+64-1 rows per warp multiplied on tensor cores and skewed through a per-warp shared-memory scratch.
// (tcgen05/TMEM version of this kernel is the next optimisation step; see DESIGN.md.)
#include "attn_common.cuh"

namespace b2c {

B2C_DEFINE_SEED_SETTER(set_seed_counter_attn)


constexpr int BQ = 64;    // query rows per CTA (4 warps x 16)
constexpr int BKV = 64;   // keys per inner tile
constexpr int BAND = 80;  // 16 + 64 - 1 rounded up to a multiple of 16
constexpr int SCR_PITCH = 84;

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}

// ---- tile loaders -------------------------------------------------------------------------------------------
// rows [row0, row0+nrows) x DK columns of a [*, ld] bf16 matrix into smem (pitch DK+8); rows outside [0, limit) → 0
template <int DK>
__device__ __forceinline__ void load_tile(bf16* s, const bf16* g, long long row0, int nrows, long long limit, int ld,
                                          int tid, int nthreads) {
  constexpr int P = DK + 8;
  constexpr int VPR = DK / 8;  // 16-byte vectors per row
  for (int i = tid; i < nrows * VPR; i += nthreads) {
    const int r = i / VPR, v = i - r * VPR;
    const long long gr = row0 + r;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (gr >= 0 && gr < limit) val = *reinterpret_cast<const uint4*>(g + gr * ld + v * 8);
    *reinterpret_cast<uint4*>(s + r * P + v * 8) = val;
  }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 → the 16 destination bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// asynchronous variant of load_tile (LDGSTS: global → smem without staging registers); rows outside [0, limit) → 0
template <int DK>
__device__ __forceinline__ void load_tile_async(bf16* s, const bf16* g, long long row0, int nrows, long long limit, int ld,
                                                int tid, int nthreads) {
  constexpr int P = DK + 8;
  constexpr int VPR = DK / 8;
  for (int i = tid; i < nrows * VPR; i += nthreads) {
    const int r = i / VPR, v = i - r * VPR;
    const long long gr = row0 + r;
    const bool ok = gr >= 0 && gr < limit;
    cp_async16(s + r * P + v * 8, ok ? g + gr * ld + v * 8 : g, ok);
  }
}

// A fragments (16 rows x DK) of bf16(q + bias) for this warp, straight from global.  rows >= T → 0
template <int DK>
__device__ __forceinline__ void load_q_frags(uint32_t (&fr)[DK / 16][4], const bf16* q, int ld, int row0, int T,
                                             const float* bias, int lane) {
  const int g = lane >> 2, c = lane & 3;
#pragma unroll
  for (int ks = 0; ks < DK / 16; ++ks) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {    // k offset 0 / 8
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {        // row g / g+8
        const int row = row0 + g + rr * 8;
        const int col = ks * 16 + half * 8 + 2 * c;
        uint32_t out = 0u;
        if (row < T) {
          float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(q + (size_t)row * ld + col));
          if (bias) { v.x += bias[col]; v.y += bias[col + 1]; }
          out = pack_bf16(v.x, v.y);
        }
        fr[ks][half * 2 + rr] = out;
      }
    }
  }
}

// S(16 x 8*NT) (+)= A(16 x DK) * Bt where Bt rows (keys / band rows) live in smem [row][DK+8].
// One ldmatrix.x4 fetches the B fragments of two adjacent 8-column tiles for one 16-wide k step.
template <int DK, int NT>
__device__ __forceinline__ void mma_qk(float (&acc)[NT][4], const uint32_t (&a)[DK / 16][4], const bf16* sB, int lane) {
  constexpr int P = DK + 8;
  static_assert(NT % 2 == 0, "NT must be even");
  const int mrow = ((lane >> 4) << 3) + (lane & 7);  // matrices 2,3 → second n-tile
  const int mcol = ((lane >> 3) & 1) << 3;           // matrices 1,3 → k offset 8
#pragma unroll
  for (int nt = 0; nt < NT; nt += 2) {
#pragma unroll
    for (int ks = 0; ks < DK / 16; ++ks) {
      uint32_t b[4];
      ldsm_x4(b, sB + (nt * 8 + mrow) * P + ks * 16 + mcol);
      mma16816(acc[nt], a[ks], b[0], b[1]);
      mma16816(acc[nt + 1], a[ks], b[2], b[3]);
    }
  }
}

// O(16 x DK) += A(16 x KROWS) * B where B rows (k) live in smem [k][DK+8]; A given as bf16 fragments per 16-k step
template <int DK, int KSTEPS>
__device__ __forceinline__ void mma_pv(float (&o)[DK / 8][4], const uint32_t (&a)[KSTEPS][4], const bf16* sB, int lane) {
  constexpr int P = DK + 8;
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
    for (int nt = 0; nt < DK / 8; ++nt) {
      uint32_t b0, b1;
      ldsm_x2_trans(b0, b1, sB + (ks * 16 + (lane & 15)) * P + nt * 8);
      mma16816(o[nt], a[ks], b0, b1);
    }
  }
}

// Recompute one 16x64 score tile for this warp: raw = ac + skew(bd), (NOT yet scaled)
template <int DK>
__device__ __forceinline__ void score_tile(float (&s)[8][4], const uint32_t (&qu)[DK / 16][4],
                                           const uint32_t (&qv)[DK / 16][4], const bf16* sK, const bf16* sP, float* scr,
                                           int warp, int lane) {
  const int g = lane >> 2, c = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
  mma_qk<DK, 8>(s, qu, sK, lane);
  float r[BAND / 8][4];
#pragma unroll
  for (int nt = 0; nt < BAND / 8; ++nt) r[nt][0] = r[nt][1] = r[nt][2] = r[nt][3] = 0.f;
  mma_qk<DK, BAND / 8>(r, qv, sP + (3 - warp) * 16 * (DK + 8), lane);
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < BAND / 8; ++nt) {
    *reinterpret_cast<float2*>(scr + g * SCR_PITCH + nt * 8 + 2 * c) = make_float2(r[nt][0], r[nt][1]);
    *reinterpret_cast<float2*>(scr + (g + 8) * SCR_PITCH + nt * 8 + 2 * c) = make_float2(r[nt][2], r[nt][3]);
  }
  __syncwarp();
  // bd[il][jl] = R[il][jl + 15 - il]
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = nt * 8 + 2 * c;
    s[nt][0] += scr[g * SCR_PITCH + col + 15 - g];
    s[nt][1] += scr[g * SCR_PITCH + col + 1 + 15 - g];
    s[nt][2] += scr[(g + 8) * SCR_PITCH + col + 15 - (g + 8)];
    s[nt][3] += scr[(g + 8) * SCR_PITCH + col + 1 + 15 - (g + 8)];
  }
}

__device__ __forceinline__ bool attn_keep(const AttnArgs& a, int b, int h, int i, int j) {
  // row pitch rounded up to even so the (j, j+1) pairs a thread owns share one hash
  const uint64_t idx = ((uint64_t)((b * a.H + h) * (int64_t)a.T + i)) * (uint64_t)((a.T + 1) & ~1) + (uint64_t)j;
  return dropout_keep(a.seed, a.site, idx, a.drop_keep);
}

// =============================================================================================== forward
template <int DK>
__global__ void __launch_bounds__(128)
relpos_attn_fwd_kernel(const AttnArgs a, bf16* __restrict__ ctx, float* __restrict__ ctx32, float* __restrict__ lse_out) {
  constexpr int P = DK + 8;
  extern __shared__ __align__(16) uint8_t smraw[];
  constexpr int STAGE = (BKV + BKV + 128) * P;  // K, V, 128 band rows of P
  bf16* stage0 = reinterpret_cast<bf16*>(smraw);
  float* scr_all = reinterpret_cast<float*>(stage0 + 2 * STAGE);  // 4 x 16 x SCR_PITCH
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;
  const int i0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int T = a.T, ld = 3 * a.d;
  const int len = a.lengths ? min(a.lengths[b], T) : T;
  const bf16* qb = a.qkv + (size_t)b * T * ld + h * DK;
  const bf16* kb = qb + a.d;
  const bf16* vb = qb + 2 * a.d;
  const bf16* pb = a.pos + h * DK;
  float* scr = scr_all + warp * 16 * SCR_PITCH;

  uint32_t qu[DK / 16][4], qv[DK / 16][4];
  load_q_frags<DK>(qu, qb, ld, i0 + warp * 16, T, a.bias_u + h * DK, lane);
  load_q_frags<DK>(qv, qb, ld, i0 + warp * 16, T, a.bias_v + h * DK, lane);

  float o[DK / 8][4];
#pragma unroll
  for (int nt = 0; nt < DK / 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  float m[2] = {-1e30f, -1e30f}, l[2] = {0.f, 0.f};
  const int row_g[2] = {i0 + warp * 16 + g, i0 + warp * 16 + g + 8};
  const float sc2 = a.scale * kLog2e;
  const uint64_t rbase[2] = {attn_row_base(a, b, h, row_g[0]), attn_row_base(a, b, h, row_g[1])};

  const int n_kv = (len + BKV - 1) / BKV;  // keys >= len contribute exactly 0 (exp(-1e4 - m) == 0 in fp32)
  auto issue = [&](int jt) {  // 2-stage cp.async ring: tile jt+1 streams in while tile jt is on the tensor cores
    bf16* st = stage0 + (jt & 1) * STAGE;
    const int j0 = jt * BKV;
    load_tile_async<DK>(st, kb, j0, BKV, T, ld, tid, 128);
    load_tile_async<DK>(st + BKV * P, vb, j0, BKV, T, ld, tid, 128);
    load_tile_async<DK>(st + 2 * BKV * P, pb, (long long)T - 1 - (i0 + 63) + j0, 128, 2LL * T - 1, a.d, tid, 128);
    cp_async_commit();
  };
  if (n_kv > 0) issue(0);
  for (int jt = 0; jt < n_kv; ++jt) {
    const int j0 = jt * BKV;
    if (jt + 1 < n_kv) { issue(jt + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const bf16* sK = stage0 + (jt & 1) * STAGE;
    const bf16* sV = sK + BKV * P;
    const bf16* sP = sV + BKV * P;
    float s[8][4];
    score_tile<DK>(s, qu, qv, sK, sP, scr, warp, lane);
    // mask + online softmax (exp2 domain)
    float mx[2] = {-1e30f, -1e30f};
    const bool rvalid[2] = {row_g[0] < len, row_g[1] < len};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int jb = j0 + nt * 8 + 2 * c;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool valid = (jb + (e & 1) < len) & rvalid[e >> 1];
        s[nt][e] = valid ? s[nt][e] * sc2 : -1e30f;
        mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
      }
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      mx[rr] = fmaxf(mx[rr], __shfl_xor_sync(0xffffffffu, mx[rr], 1));
      mx[rr] = fmaxf(mx[rr], __shfl_xor_sync(0xffffffffu, mx[rr], 2));
    }
    float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const float mn = fmaxf(m[rr], mx[rr]);
      corr[rr] = exp2f(m[rr] - mn);
      m[rr] = mn;
    }
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float p[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        p[e] = exp2f(s[nt][e] - m[e >> 1]);     // masked entries: exp2(-1e30 - m) == 0 (m is finite for valid rows)
        p[e] = (s[nt][e] > -1e29f) ? p[e] : 0.f;  // fully-masked rows (m == -1e30)
        rs[e >> 1] += p[e];
      }
      if (a.drop_keep != 0) {
        const int jb = j0 + nt * 8 + 2 * c;
        const uint32_t k0 = attn_keep2(a, rbase[0], jb), k1 = attn_keep2(a, rbase[1], jb);
        p[0] = (k0 & 1u) ? p[0] * a.drop_scale : 0.f;
        p[1] = (k0 & 2u) ? p[1] * a.drop_scale : 0.f;
        p[2] = (k1 & 1u) ? p[2] * a.drop_scale : 0.f;
        p[3] = (k1 & 2u) ? p[3] * a.drop_scale : 0.f;
      }
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p[0], p[1]);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p[2], p[3]);
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      rs[rr] += __shfl_xor_sync(0xffffffffu, rs[rr], 1);
      rs[rr] += __shfl_xor_sync(0xffffffffu, rs[rr], 2);
      l[rr] = l[rr] * corr[rr] + rs[rr];
    }
#pragma unroll
    for (int nt = 0; nt < DK / 8; ++nt) {
      o[nt][0] *= corr[0]; o[nt][1] *= corr[0];
      o[nt][2] *= corr[1]; o[nt][3] *= corr[1];
    }
    mma_pv<DK, 4>(o, pa, sV, lane);
    __syncthreads();  // every warp is done with this stage before the next issue() overwrites it
  }
  // finalise: rows >= len (fully masked queries) output exactly 0 (second masked_fill, :137-138)
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int row = row_g[rr];
    if (row < T) {
      const float inv = (row < len && l[rr] > 0.f) ? 1.f / l[rr] : 0.f;
      bf16* op = ctx + ((size_t)b * T + row) * a.d + h * DK;
#pragma unroll
      for (int nt = 0; nt < DK / 8; ++nt)
        *reinterpret_cast<uint32_t*>(op + nt * 8 + 2 * c) = pack_bf16(o[nt][rr * 2] * inv, o[nt][rr * 2 + 1] * inv);
      if (ctx32) {  // un-rounded copy for the backward's delta = rowsum(dO * O) (a bf16 O biases whole rows of dS)
        float* o32 = ctx32 + ((size_t)b * T + row) * a.d + h * DK;
#pragma unroll
        for (int nt = 0; nt < DK / 8; ++nt)
          *reinterpret_cast<float2*>(o32 + nt * 8 + 2 * c) = make_float2(o[nt][rr * 2] * inv, o[nt][rr * 2 + 1] * inv);
      }
      if (c == 0 && lse_out) lse_out[((size_t)b * a.H + h) * T + row] = (row < len && l[rr] > 0.f) ? m[rr] + log2f(l[rr]) : 0.f;
    }
  }
}

// =============================================================================================== backward (query side)
// CTA owns 64 query rows: recomputes S/P, dP = dO V^T, dS; accumulates dq = dS K + unskew(dS) Pband, d(bias_u/v);
// stages dS (scaled, bf16) and P_dropped (bf16) in HBM workspaces [B,H,T,Tp] for the key-side / position-side kernels.
template <int DK>
__global__ void __launch_bounds__(128)
relpos_attn_bwd_q_kernel(const AttnArgs a, const bf16* __restrict__ ctx, const float* __restrict__ ctx32,
                         const bf16* __restrict__ dctx,
                         const float* __restrict__ lse, bf16* __restrict__ dqkv, float* __restrict__ dbias_u,
                         float* __restrict__ dbias_v, bf16* __restrict__ ws_ds, bf16* __restrict__ ws_p, int Tp) {
  constexpr int P = DK + 8;
  extern __shared__ __align__(16) uint8_t smraw[];
  constexpr int STAGE = (BKV + BKV + 128) * P;
  bf16* stage0 = reinterpret_cast<bf16*>(smraw);
  float* scr_all = reinterpret_cast<float*>(stage0 + 2 * STAGE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;
  const int i0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int T = a.T, ld = 3 * a.d;
  const int len = a.lengths ? min(a.lengths[b], T) : T;
  const bf16* qb = a.qkv + (size_t)b * T * ld + h * DK;
  const bf16* kb = qb + a.d;
  const bf16* vb = qb + 2 * a.d;
  const bf16* pb = a.pos + h * DK;
  float* scr = scr_all + warp * 16 * SCR_PITCH;
  const int wrow0 = i0 + warp * 16;
  const int row_g[2] = {wrow0 + g, wrow0 + g + 8};

  uint32_t qu[DK / 16][4], qv[DK / 16][4], dof[DK / 16][4];
  load_q_frags<DK>(qu, qb, ld, wrow0, T, a.bias_u + h * DK, lane);
  load_q_frags<DK>(qv, qb, ld, wrow0, T, a.bias_v + h * DK, lane);
  load_q_frags<DK>(dof, dctx + (size_t)b * T * a.d + h * DK, a.d, wrow0, T, nullptr, lane);

  // delta_i = sum_d dO[i,d] * O[i,d]
  float delta[2] = {0.f, 0.f}, lse2[2] = {0.f, 0.f};
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int row = row_g[rr];
    float acc = 0.f;
    if (row < T) {
      const bf16* op = ctx + ((size_t)b * T + row) * a.d + h * DK;
      const bf16* dp = dctx + ((size_t)b * T + row) * a.d + h * DK;
      const float* o32 = ctx32 ? ctx32 + ((size_t)b * T + row) * a.d + h * DK : nullptr;
#pragma unroll
      for (int nt = 0; nt < DK / 8; ++nt) {
        const float2 x = o32 ? *reinterpret_cast<const float2*>(o32 + nt * 8 + 2 * c)
                             : unpack_bf16(*reinterpret_cast<const uint32_t*>(op + nt * 8 + 2 * c));
        const float2 y = unpack_bf16(*reinterpret_cast<const uint32_t*>(dp + nt * 8 + 2 * c));
        acc += x.x * y.x + x.y * y.y;
      }
      lse2[rr] = lse[((size_t)b * a.H + h) * T + row];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);  // shuffles outside the divergent branch
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    delta[rr] = acc;
  }

  float dqu[DK / 8][4], dqv[DK / 8][4];
#pragma unroll
  for (int nt = 0; nt < DK / 8; ++nt) {
    dqu[nt][0] = dqu[nt][1] = dqu[nt][2] = dqu[nt][3] = 0.f;
    dqv[nt][0] = dqv[nt][1] = dqv[nt][2] = dqv[nt][3] = 0.f;
  }
  const float sc2 = a.scale * kLog2e;
  const uint64_t rbase[2] = {attn_row_base(a, b, h, row_g[0]), attn_row_base(a, b, h, row_g[1])};
  const bool rvalid[2] = {row_g[0] < len, row_g[1] < len};
  bf16* wds = ws_ds + ((size_t)(b * a.H + h) * T) * Tp;
  bf16* wp = ws_p + ((size_t)(b * a.H + h) * T) * Tp;
  const int n_kv_all = (T + BKV - 1) / BKV;
  const int n_kv = (len + BKV - 1) / BKV;
  auto issue = [&](int jt) {
    bf16* st = stage0 + (jt & 1) * STAGE;
    const int j0 = jt * BKV;
    load_tile_async<DK>(st, kb, j0, BKV, T, ld, tid, 128);
    load_tile_async<DK>(st + BKV * P, vb, j0, BKV, T, ld, tid, 128);
    load_tile_async<DK>(st + 2 * BKV * P, pb, (long long)T - 1 - (i0 + 63) + j0, 128, 2LL * T - 1, a.d, tid, 128);
    cp_async_commit();
  };
  if (n_kv > 0) issue(0);
  for (int jt = 0; jt < n_kv_all; ++jt) {
    const int j0 = jt * BKV;
    if (jt >= n_kv) {  // tiles fully beyond the valid length: dS = P = 0, just clear the workspaces
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int j = j0 + nt * 8 + 2 * c;
          if (row_g[rr] < T && j < Tp) {
            *reinterpret_cast<uint32_t*>(wds + (size_t)row_g[rr] * Tp + j) = 0u;
            *reinterpret_cast<uint32_t*>(wp + (size_t)row_g[rr] * Tp + j) = 0u;
          }
        }
      continue;
    }
    if (jt + 1 < n_kv) { issue(jt + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const bf16* sK = stage0 + (jt & 1) * STAGE;
    const bf16* sV = sK + BKV * P;
    const bf16* sP = sV + BKV * P;
    float s[8][4];
    score_tile<DK>(s, qu, qv, sK, sP, scr, warp, lane);
    float dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
    mma_qk<DK, 8>(dp, dof, sV, lane);  // dP[i,j] = dO_i . V_j
    uint32_t dsa[4][4];                // dS as A fragments (keys = k dimension)
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float ds[4], pd[4];
      const int jb = j0 + nt * 8 + 2 * c;
      uint32_t keep[2] = {3u, 3u};
      if (a.drop_keep != 0) {
        keep[0] = attn_keep2(a, rbase[0], jb);
        keep[1] = attn_keep2(a, rbase[1], jb);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int rr = e >> 1;
        const bool valid = (jb + (e & 1) < len) & rvalid[rr];
        const float p = valid ? exp2f(s[nt][e] * sc2 - lse2[rr]) : 0.f;
        const bool kp = (keep[rr] >> (e & 1)) & 1u;
        pd[e] = kp ? p * a.drop_scale : 0.f;
        const float dpe = kp ? dp[nt][e] * a.drop_scale : 0.f;
        ds[e] = p * (dpe - delta[rr]) * a.scale;
      }
      dsa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(ds[0], ds[1]);
      dsa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
      // stage for the key-side / position-side kernels
      const int j = j0 + nt * 8 + 2 * c;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        if (row_g[rr] < T && j < Tp) {
          *reinterpret_cast<uint32_t*>(wds + (size_t)row_g[rr] * Tp + j) = pack_bf16(ds[rr * 2], ds[rr * 2 + 1]);
          *reinterpret_cast<uint32_t*>(wp + (size_t)row_g[rr] * Tp + j) = pack_bf16(pd[rr * 2], pd[rr * 2 + 1]);
        }
      }
      // unskewed copy for the positional path: scr[il][jl + 15 - il] = dS[il][jl]
      scr[g * SCR_PITCH + nt * 8 + 2 * c + 15 - g] = bf16_round(ds[0]);
      scr[g * SCR_PITCH + nt * 8 + 2 * c + 1 + 15 - g] = bf16_round(ds[1]);
      scr[(g + 8) * SCR_PITCH + nt * 8 + 2 * c + 15 - (g + 8)] = bf16_round(ds[2]);
      scr[(g + 8) * SCR_PITCH + nt * 8 + 2 * c + 1 + 15 - (g + 8)] = bf16_round(ds[3]);
    }
    __syncwarp();
    // dQ_u += dS K
    mma_pv<DK, 4>(dqu, dsa, sK, lane);
    // dQ_v += unskew(dS) (16 x 80) * Pband (80 x DK)
    uint32_t ua[BAND / 16][4];
#pragma unroll
    for (int ks = 0; ks < BAND / 16; ++ks) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int il = g + rr * 8;
          const int n = ks * 16 + half * 8 + 2 * c;
          // valid band columns for row il: [15 - il, 15 - il + 64)
          const int lo = 15 - il, hi = 15 - il + 64;
          const float x0 = (n >= lo && n < hi) ? scr[il * SCR_PITCH + n] : 0.f;
          const float x1 = (n + 1 >= lo && n + 1 < hi) ? scr[il * SCR_PITCH + n + 1] : 0.f;
          ua[ks][half * 2 + rr] = pack_bf16(x0, x1);
        }
      }
    }
    mma_pv<DK, BAND / 16>(dqv, ua, sP + (3 - warp) * 16 * P, lane);
    __syncthreads();
  }
  // dq = dQ_u + dQ_v → bf16 into the q slot of dqkv
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int row = row_g[rr];
    if (row < T) {
      bf16* op = dqkv + ((size_t)b * T + row) * ld + h * DK;
#pragma unroll
      for (int nt = 0; nt < DK / 8; ++nt)
        *reinterpret_cast<uint32_t*>(op + nt * 8 + 2 * c) =
            pack_bf16(dqu[nt][rr * 2] + dqv[nt][rr * 2], dqu[nt][rr * 2 + 1] + dqv[nt][rr * 2 + 1]);
    }
  }
  // d(pos_bias_u/v)[h, col] += sum over this warp's rows
#pragma unroll
  for (int nt = 0; nt < DK / 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float su = dqu[nt][e] + dqu[nt][e + 2];
      float sv = dqv[nt][e] + dqv[nt][e + 2];
#pragma unroll
      for (int off = 4; off < 32; off <<= 1) {
        su += __shfl_xor_sync(0xffffffffu, su, off);
        sv += __shfl_xor_sync(0xffffffffu, sv, off);
      }
      if (g == 0) {
        atomicAdd(dbias_u + h * DK + nt * 8 + 2 * c + e, su);
        atomicAdd(dbias_v + h * DK + nt * 8 + 2 * c + e, sv);
      }
    }
  }
}

// =============================================================================================== backward (key side)
// CTA owns 64 keys of one (b,h): dK[j,:] = sum_i dS[i,j] * bf16(q_i+u);  dV[j,:] = sum_i Pd[i,j] * dO[i,:]
template <int DK>
__global__ void __launch_bounds__(128)
relpos_attn_bwd_kv_kernel(const AttnArgs a, const bf16* __restrict__ dctx, const bf16* __restrict__ ws_ds,
                          const bf16* __restrict__ ws_p, bf16* __restrict__ dqkv, int Tp) {
  constexpr int P = DK + 8;
  constexpr int PX = BKV + 8;
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* sQ = reinterpret_cast<bf16*>(smraw);  // [64 i][DK]  bf16(q+u)
  bf16* sO = sQ + 64 * P;                     // [64 i][DK]  dO
  bf16* sDS = sO + 64 * P;                    // [64 i][64 j]
  bf16* sPd = sDS + 64 * PX;                  // [64 i][64 j]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;
  const int j0 = blockIdx.x * BKV, h = blockIdx.y, b = blockIdx.z;
  const int T = a.T, ld = 3 * a.d;
  const int len = a.lengths ? min(a.lengths[b], T) : T;
  const bf16* qb = a.qkv + (size_t)b * T * ld + h * DK;
  const bf16* dob = dctx + (size_t)b * T * a.d + h * DK;
  const bf16* wds = ws_ds + ((size_t)(b * a.H + h) * T) * Tp;
  const bf16* wp = ws_p + ((size_t)(b * a.H + h) * T) * Tp;
  const float* bu = a.bias_u + h * DK;

  float dk_[DK / 8][4], dv_[DK / 8][4];
#pragma unroll
  for (int nt = 0; nt < DK / 8; ++nt) {
    dk_[nt][0] = dk_[nt][1] = dk_[nt][2] = dk_[nt][3] = 0.f;
    dv_[nt][0] = dv_[nt][1] = dv_[nt][2] = dv_[nt][3] = 0.f;
  }
  constexpr int STAGE = 2 * 64 * P + 2 * 64 * PX;
  const int n_q = (len + 63) / 64;  // query rows >= len have dS = P = 0
  auto issue = [&](int it) {       // 2-stage cp.async ring over the query tiles
    bf16* st = sQ + (it & 1) * STAGE;
    const int i0 = it * 64;
    load_tile_async<DK>(st, qb, i0, 64, T, ld, tid, 128);                // raw q (bias added in place below)
    load_tile_async<DK>(st + 64 * P, dob, i0, 64, T, a.d, tid, 128);     // dO
    bf16* d1 = st + 2 * 64 * P;
    bf16* d2 = d1 + 64 * PX;
    for (int idx = tid; idx < 64 * 8; idx += 128) {
      const int r = idx >> 3, v = idx & 7;
      const bool ok = (i0 + r < T) && (j0 + v * 8 < Tp);
      const size_t off = ok ? (size_t)(i0 + r) * Tp + j0 + v * 8 : 0;
      cp_async16(d1 + r * PX + v * 8, wds + off, ok);
      cp_async16(d2 + r * PX + v * 8, wp + off, ok);
    }
    cp_async_commit();
  };
  if (n_q > 0) issue(0);
  for (int it = 0; it < n_q; ++it) {
    if (it + 1 < n_q) { issue(it + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    bf16* tQ = sQ + (it & 1) * STAGE;
    const bf16* tO = tQ + 64 * P;
    const bf16* tDS = tO + 64 * P;
    const bf16* tPd = tDS + 64 * PX;
    // q → bf16(q + u) in place (each thread owns its elements; rows >= T are zero-filled and must stay zero)
    for (int idx = tid; idx < 64 * (DK / 2); idx += 128) {
      const int r = idx / (DK / 2), cp = idx - r * (DK / 2);
      if (it * 64 + r < T) {
        uint32_t* pq = reinterpret_cast<uint32_t*>(tQ + r * P + cp * 2);
        const float2 v = unpack_bf16(*pq);
        *pq = pack_bf16(v.x + bu[cp * 2], v.y + bu[cp * 2 + 1]);
      }
    }
    __syncthreads();
    // this warp's 16 keys: columns warp*16.. of dS / Pd (A = X^T via ldmatrix.trans)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ads[4], apd[4];
      const int mi = lane >> 3;
      const int krow = ks * 16 + (mi >> 1) * 8 + (lane & 7);
      const int mcol = warp * 16 + (mi & 1) * 8;
      ldsm_x4_trans(ads, tDS + krow * PX + mcol);
      ldsm_x4_trans(apd, tPd + krow * PX + mcol);
#pragma unroll
      for (int nt = 0; nt < DK / 8; ++nt) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, tQ + (ks * 16 + (lane & 15)) * P + nt * 8);
        mma16816(dk_[nt], ads, b0, b1);
        ldsm_x2_trans(b0, b1, tO + (ks * 16 + (lane & 15)) * P + nt * 8);
        mma16816(dv_[nt], apd, b0, b1);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int j = j0 + warp * 16 + g + rr * 8;
    if (j < T) {
      bf16* okp = dqkv + ((size_t)b * T + j) * ld + a.d + h * DK;
      bf16* ovp = okp + a.d;
#pragma unroll
      for (int nt = 0; nt < DK / 8; ++nt) {
        *reinterpret_cast<uint32_t*>(okp + nt * 8 + 2 * c) = pack_bf16(dk_[nt][rr * 2], dk_[nt][rr * 2 + 1]);
        *reinterpret_cast<uint32_t*>(ovp + nt * 8 + 2 * c) = pack_bf16(dv_[nt][rr * 2], dv_[nt][rr * 2 + 1]);
      }
    }
  }
}

// =============================================================================================== backward (position side)
// CTA owns 64 relative-position rows r of one head for a chunk of the batch:
//   dpos[r, h*DK + :] += sum_{b in chunk} sum_i dS[b,h,i, i-(T-1)+r] * bf16(q_{b,i}+v)      (vector reductions at the end)
template <int DK>
__global__ void __launch_bounds__(128)
relpos_attn_bwd_pos_kernel(const AttnArgs a, const bf16* __restrict__ ws_ds, float* __restrict__ dpos, int Tp,
                           int b_per_cta) {
  constexpr int P = DK + 8;
  constexpr int PX = 64 + 8;
  constexpr int QV = 64 * DK / 8 / 128;  // 16-byte chunks of the q tile per thread
  extern __shared__ __align__(16) uint8_t smraw[];
  bf16* sQ = reinterpret_cast<bf16*>(smraw);  // [64 i][DK] bf16(q+v)
  bf16* sX = sQ + 64 * P;                     // [64 i][64 r]  diagonal gather of dS
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;
  const int r0 = blockIdx.x * 64, h = blockIdx.y;
  const int T = a.T, ld = 3 * a.d, R = 2 * T - 1;
  float acc[DK / 8][4];
#pragma unroll
  for (int nt = 0; nt < DK / 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
  // rows i that can pair with r in [r0, r0+64): j = i - (T-1) + r in [0, T)  →  i in [T-1-r0-63, 2T-2-r0]
  const int i_lo = max(0, T - 1 - (r0 + 63)), i_hi = min(T - 1, 2 * T - 2 - r0);
  const int i_first = (i_lo / 64) * 64;
  const int b_lo = blockIdx.z * b_per_cta, b_hi = min(a.B, b_lo + b_per_cta);
  // this thread's slice of the two tiles: q chunk column (fixed) and the gather row / half
  const int qck = tid & (DK / 8 - 1);
  float bvr[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) bvr[k] = a.bias_v[h * DK + qck * 8 + k];
  const int il = tid >> 1, half = tid & 1;

  // The (batch, 64-row block) iterations are software-pipelined: the global loads of step n+1 (q rows as 16-byte
  // chunks; the diagonal of dS as five ALIGNED 16-byte loads per thread, realigned in registers by a word barrel
  // shift + half-word funnel shift) are in flight while the tensor cores work on step n.
  struct Step { int b, i0, hi; };
  auto first_step = [&](int b) -> Step {
    Step st{b, i_first, -1};
    while (st.b < b_hi) {
      const int len = a.lengths ? min(a.lengths[st.b], T) : T;
      st.hi = min(i_hi, len - 1);
      if (st.i0 <= st.hi) return st;
      ++st.b;
    }
    return st;
  };
  auto next_step = [&](Step st) -> Step {
    st.i0 += 64;
    if (st.i0 <= st.hi) return st;
    return first_step(st.b + 1);
  };
  uint4 qreg[QV];
  uint4 greg[5];
  int gmode = 0;  // 0 zeros, 1 aligned fast path (shift = gsh), 2 edge row (element-wise at store time)
  int gsh = 0;
  auto prefetch = [&](const Step& st) {
    const bf16* qb = a.qkv + (size_t)st.b * T * ld + h * DK;
#pragma unroll
    for (int k = 0; k < QV; ++k) {
      const int r = (tid + 128 * k) / (DK / 8);
      qreg[k] = make_uint4(0, 0, 0, 0);
      if (st.i0 + r < T) qreg[k] = *reinterpret_cast<const uint4*>(qb + (size_t)(st.i0 + r) * ld + qck * 8);
    }
    const int i = st.i0 + il;
    const int js = i - (T - 1) + r0 + half * 32;
    if (i > st.hi) {
      gmode = 0;
    } else if (js >= 0 && js + 32 <= Tp) {  // columns [T, Tp) of the staged dS are zeros
      gmode = 1;
      gsh = js & 7;
      const uint4* src = reinterpret_cast<const uint4*>(ws_ds + ((size_t)(st.b * a.H + h) * T + i) * Tp + (js & ~7));
#pragma unroll
      for (int k = 0; k < 5; ++k) greg[k] = __ldg(src + k);
    } else {
      gmode = 2;
    }
  };
  auto publish = [&](const Step& st) {
#pragma unroll
    for (int k = 0; k < QV; ++k) {
      const int r = (tid + 128 * k) / (DK / 8);
      const uint32_t w[4] = {qreg[k].x, qreg[k].y, qreg[k].z, qreg[k].w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16(w[e]);
        o[e] = (st.i0 + r < T) ? pack_bf16(f.x + bvr[2 * e], f.y + bvr[2 * e + 1]) : 0u;
      }
      *reinterpret_cast<uint4*>(sQ + r * P + qck * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    uint4* dst = reinterpret_cast<uint4*>(sX + il * PX + half * 32);
    if (gmode == 1) {
      uint32_t w[21];
#pragma unroll
      for (int k = 0; k < 5; ++k) { w[4 * k] = greg[k].x; w[4 * k + 1] = greg[k].y; w[4 * k + 2] = greg[k].z; w[4 * k + 3] = greg[k].w; }
      w[20] = 0u;
      const uint32_t wsft = (uint32_t)gsh >> 1, odd16 = (uint32_t)(gsh & 1) * 16u;
#pragma unroll
      for (int k = 0; k < 18; ++k) w[k] = (wsft & 2u) ? w[k + 2] : w[k];
#pragma unroll
      for (int k = 0; k < 17; ++k) w[k] = (wsft & 1u) ? w[k + 1] : w[k];
#pragma unroll
      for (int k = 0; k < 16; ++k) w[k] = __funnelshift_r(w[k], w[k + 1], odd16);
#pragma unroll
      for (int k = 0; k < 4; ++k) dst[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    } else if (gmode == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) dst[k] = make_uint4(0, 0, 0, 0);
    } else {
      const int i = st.i0 + il;
      const int js = i - (T - 1) + r0 + half * 32;
      const bf16* wrow = ws_ds + ((size_t)(st.b * a.H + h) * T + i) * Tp;
      bf16* d16 = reinterpret_cast<bf16*>(dst);
      for (int k = 0; k < 32; ++k) {
        const int j = js + k;
        d16[k] = (j >= 0 && j < T) ? wrow[j] : __float2bfloat16_rn(0.f);
      }
    }
  };

  Step cur = first_step(b_lo);
  if (cur.b < b_hi) prefetch(cur);
  while (cur.b < b_hi) {
    __syncthreads();  // the previous step's MMAs have read the tiles
    publish(cur);
    const Step nxt = next_step(cur);
    if (nxt.b < b_hi) prefetch(nxt);
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t ax[4];
      const int mi = lane >> 3;
      ldsm_x4_trans(ax, sX + (ks * 16 + (mi >> 1) * 8 + (lane & 7)) * PX + warp * 16 + (mi & 1) * 8);
#pragma unroll
      for (int nt = 0; nt < DK / 8; ++nt) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, sQ + (ks * 16 + (lane & 15)) * P + nt * 8);
        mma16816(acc[nt], ax, b0, b1);
      }
    }
    cur = nxt;
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int r = r0 + warp * 16 + g + rr * 8;
    if (r < R) {
      float* op = dpos + (size_t)r * a.d + h * DK;
#pragma unroll
      for (int nt = 0; nt < DK / 8; ++nt) red_add_v2(op + nt * 8 + 2 * c, acc[nt][rr * 2], acc[nt][rr * 2 + 1]);
    }
  }
}

static int fill_args(AttnArgs& a, const void* qkv, const void* pos, const float* bu, const float* bv,
                     const int32_t* lengths, int B, int T, int h, int dk, float drop_p, unsigned long long seed,
                     unsigned int site) {
  B2C_REQUIRE(dk == 64 || dk == 128, "relpos_attn: head dim %d not built (64, 128)", dk);
  B2C_REQUIRE(B > 0 && T > 0 && h > 0, "relpos_attn: bad shape");
  a.qkv = reinterpret_cast<const bf16*>(qkv);
  a.pos = reinterpret_cast<const bf16*>(pos);
  a.bias_u = bu; a.bias_v = bv; a.lengths = lengths;
  a.B = B; a.T = T; a.H = h; a.d = h * dk;
  a.scale = 1.0f / sqrtf((float)dk);
  a.drop_p = drop_p; a.seed = seed; a.site = site;
  a.drop_keep = 0; a.drop_scale = 1.f;
  if (drop_p > 0.f) {
    a.drop_keep = (unsigned int)((1.0 - (double)drop_p) * 65536.0 + 0.5);
    a.drop_scale = 1.f / (1.f - drop_p);
  }
  return B2C_OK;
}

template <int DK>
static size_t smem_q() { return (size_t)2 * (64 + 64 + 128) * (DK + 8) * 2 + 4 * 16 * SCR_PITCH * 4; }

template <int DK>
static int launch_fwd(const AttnArgs& a, bf16* ctx, float* ctx32, float* lse, cudaStream_t st) {
  const size_t smem = smem_q<DK>();
  static bool set = false;
  if (!set) {
    B2C_CUDA(cudaFuncSetAttribute(relpos_attn_fwd_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    set = true;
  }
  relpos_attn_fwd_kernel<DK><<<dim3((a.T + BQ - 1) / BQ, a.H, a.B), 128, smem, st>>>(a, ctx, ctx32, lse);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

template <int DK>
static int launch_bwd(const AttnArgs& a, const bf16* ctx, const float* ctx32, const bf16* dctx, const float* lse, bf16* dqkv, float* dpos,
                      float* dbu, float* dbv, bf16* ws_ds, bf16* ws_p, bf16* ws_qu, int Tp, cudaStream_t st) {
  const size_t smem1 = smem_q<DK>();
  const size_t smem2 = 2 * ((size_t)2 * 64 * (DK + 8) * 2 + (size_t)2 * 64 * (BKV + 8) * 2);
  const size_t smem3 = (size_t)64 * (DK + 8) * 2 + (size_t)64 * 72 * 2;
  static bool set = false;
  if (!set) {
    B2C_CUDA(cudaFuncSetAttribute(relpos_attn_bwd_q_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    B2C_CUDA(cudaFuncSetAttribute(relpos_attn_bwd_kv_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    B2C_CUDA(cudaFuncSetAttribute(relpos_attn_bwd_pos_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
    set = true;
  }
  const int qt = (a.T + BQ - 1) / BQ;
  // B2C_ATTN_IMPL=mma selects the legacy mma.sync kernels, =mmakv only the legacy key-side kernel (A/B comparison)
  static int impl = -1;
  if (impl < 0) {
    const char* e = getenv("B2C_ATTN_IMPL");
    impl = (e && !strcmp(e, "mma")) ? 1 : (e && !strcmp(e, "mmakv")) ? 2 : 0;
  }
  if (DK == 64 && impl != 1) {
    int rc = launch_relpos_attn_bwd_q_tc(a, ctx, ctx32, dctx, lse, dqkv, dbu, dbv, ws_ds, ws_p, ws_qu, Tp, st);
    if (rc) return rc;
  } else {
    relpos_attn_bwd_q_kernel<DK><<<dim3(qt, a.H, a.B), 128, smem1, st>>>(a, ctx, ctx32, dctx, lse, dqkv, dbu, dbv, ws_ds, ws_p, Tp);
    B2C_LAUNCH_CHECK();
  }
  if (DK == 64 && impl == 0) {
    int rc = launch_relpos_attn_bwd_kv_tc(a, dctx, ws_ds, ws_p, ws_qu, dqkv, Tp, st);
    if (rc) return rc;
  } else {
    relpos_attn_bwd_kv_kernel<DK><<<dim3(qt, a.H, a.B), 128, smem2, st>>>(a, dctx, ws_ds, ws_p, dqkv, Tp);
    B2C_LAUNCH_CHECK();
  }
  const int b_per_cta = 2;
  relpos_attn_bwd_pos_kernel<DK><<<dim3((2 * a.T - 1 + 63) / 64, a.H, (a.B + b_per_cta - 1) / b_per_cta), 128, smem3, st>>>(
      a, ws_ds, dpos, Tp, b_per_cta);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_relpos_attn_fwd(const void* qkv, const void* pos, const float* bias_u, const float* bias_v,
                                   const int32_t* lengths, void* ctx, float* ctx_f32, float* lse, int B, int T, int h,
                                   int dk, float drop_p, unsigned long long seed, unsigned int site, void* stream) {
  AttnArgs a;
  int rc = fill_args(a, qkv, pos, bias_u, bias_v, lengths, B, T, h, dk, drop_p, seed, site);
  if (rc) return rc;
  if (dk == 64) {
    static int impl = -1;  // B2C_ATTN_IMPL=mma selects the legacy mma.sync kernel (debug / A-B comparison)
    if (impl < 0) {
      const char* e = getenv("B2C_ATTN_IMPL");
      impl = (e && !strcmp(e, "mma")) ? 1 : 0;
    }
    if (impl == 0) return launch_relpos_attn_fwd_tc(a, reinterpret_cast<bf16*>(ctx), ctx_f32, lse, (cudaStream_t)stream);
    return launch_fwd<64>(a, reinterpret_cast<bf16*>(ctx), ctx_f32, lse, (cudaStream_t)stream);
  }
  return launch_fwd<128>(a, reinterpret_cast<bf16*>(ctx), ctx_f32, lse, (cudaStream_t)stream);
}

/* workspace: dS and P_dropped, 2 * B*h*T*Tp bf16 with Tp = (T+7)&~7, then (q + pos_bias_u) [B,T,h*dk] bf16 (dk <= 128) */
extern "C" size_t b2c_relpos_attn_bwd_workspace(int B, int T, int h) {
  const size_t Tp = (size_t)((T + 7) & ~7);
  return 2 * (size_t)B * h * T * Tp * 2 + (size_t)B * T * h * 128 * 2;
}

extern "C" int b2c_relpos_attn_bwd(const void* qkv, const void* pos, const float* bias_u, const float* bias_v,
                                   const int32_t* lengths, const void* ctx, const float* ctx_f32, const void* dctx,
                                   const float* lse, void* dqkv, float* dpos_f32, float* dbias_u, float* dbias_v, void* workspace,
                                   size_t workspace_bytes, int B, int T, int h, int dk, float drop_p,
                                   unsigned long long seed, unsigned int site, void* stream) {
  AttnArgs a;
  int rc = fill_args(a, qkv, pos, bias_u, bias_v, lengths, B, T, h, dk, drop_p, seed, site);
  if (rc) return rc;
  const int Tp = (T + 7) & ~7;
  B2C_REQUIRE(workspace && workspace_bytes >= b2c_relpos_attn_bwd_workspace(B, T, h), "relpos_attn_bwd: workspace too small");
  bf16* ws_ds = reinterpret_cast<bf16*>(workspace);
  bf16* ws_p = ws_ds + (size_t)B * h * T * Tp;
  bf16* ws_qu = ws_p + (size_t)B * h * T * Tp;
  if (dk == 64)
    return launch_bwd<64>(a, reinterpret_cast<const bf16*>(ctx), ctx_f32, reinterpret_cast<const bf16*>(dctx), lse,
                          reinterpret_cast<bf16*>(dqkv), dpos_f32, dbias_u, dbias_v, ws_ds, ws_p, ws_qu, Tp, (cudaStream_t)stream);
  return launch_bwd<128>(a, reinterpret_cast<const bf16*>(ctx), ctx_f32, reinterpret_cast<const bf16*>(dctx), lse,
                         reinterpret_cast<bf16*>(dqkv), dpos_f32, dbias_u, dbias_v, ws_ds, ws_p, ws_qu, Tp, (cudaStream_t)stream);
}
