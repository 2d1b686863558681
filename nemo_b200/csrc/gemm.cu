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

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kNumThreads = 256;

struct GemmParams {
  int M, N, K;
  int splits;       // split-K factor (fp32 atomic accumulate epilogue only)
  int k_per_split;  // in units of BK blocks
  b2c_epilogue epi;
};

template <int BN>
struct SmemLayout {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // + barriers + alignment slack
};

// ---------------------------------------------------------------- epilogue
// Each call handles one accumulator row `m`, 32 consecutive columns starting at n0.
template <int MODE>
__device__ __forceinline__ void epilogue_row(const b2c_epilogue& e, int m, int n0, int M, int N, const uint32_t (&acc)[32],
                                             bool atomic) {
  if (m >= M || n0 >= N) return;
  const bool full = (n0 + 32 <= N);
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  if (MODE != B2C_EPI_ACCUM_F32 && e.bias != nullptr) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b = *reinterpret_cast<const float4*>(e.bias + n0 + j);
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
      for (int j = 0; j < 32 && n0 + j < N; ++j) v[j] += e.bias[n0 + j];
    }
  }
  const size_t ldo = (size_t)e.ldo;

  if (MODE == B2C_EPI_BF16) {
    // out_bf16[m, n] = bf16(acc + bias)
    bf16* o = reinterpret_cast<bf16*>(e.out0) + (size_t)m * ldo + n0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 p = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]), pack_bf16(v[j + 4], v[j + 5]),
                             pack_bf16(v[j + 6], v[j + 7]));
        *reinterpret_cast<uint4*>(o + j) = p;
      }
    } else {
      for (int j = 0; j < 32 && n0 + j < N; ++j) o[j] = __float2bfloat16_rn(v[j]);
    }
  } else if (MODE == B2C_EPI_SILU) {
    // z = bf16(acc+bias) → out0 ; h = bf16(silu(z)) [* dropout] → out1   (conformer_modules.py:383-385)
    bf16* oz = reinterpret_cast<bf16*>(e.out0) + (size_t)m * ldo + n0;
    bf16* oh = reinterpret_cast<bf16*>(e.out1) + (size_t)m * ldo + n0;
    float h[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float z = bf16_round(v[j]);
      v[j] = z;
      h[j] = silu_f(z);
    }
    if (e.drop_keep != 0) {
      const float inv = e.drop_scale;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        uint64_t idx = (uint64_t)m * (uint64_t)N + (uint64_t)(n0 + j);
        h[j] = dropout_keep(e.seed, e.drop_site, idx, e.drop_keep) ? bf16_round(h[j]) * inv : 0.f;
      }
    }
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        *reinterpret_cast<uint4*>(oz + j) = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                                                       pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
        *reinterpret_cast<uint4*>(oh + j) = make_uint4(pack_bf16(h[j], h[j + 1]), pack_bf16(h[j + 2], h[j + 3]),
                                                       pack_bf16(h[j + 4], h[j + 5]), pack_bf16(h[j + 6], h[j + 7]));
      }
    } else {
      for (int j = 0; j < 32 && n0 + j < N; ++j) {
        oz[j] = __float2bfloat16_rn(v[j]);
        oh[j] = __float2bfloat16_rn(h[j]);
      }
    }
  } else if (MODE == B2C_EPI_RESID_F32) {
    // out0_f32[m,n] = (aux_f32 ? aux_f32[m,n] : 0) + alpha * drop(bf16(acc+bias))
    float* o = reinterpret_cast<float*>(e.out0) + (size_t)m * ldo + n0;
    const float* r = e.aux ? reinterpret_cast<const float*>(e.aux) + (size_t)m * ldo + n0 : nullptr;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);
    if (e.drop_keep != 0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        uint64_t idx = (uint64_t)m * (uint64_t)N + (uint64_t)(n0 + j);
        v[j] = dropout_keep(e.seed, e.drop_site, idx, e.drop_keep) ? v[j] * e.drop_scale : 0.f;
      }
    }
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 rr = r ? *reinterpret_cast<const float4*>(r + j) : make_float4(0, 0, 0, 0);
        float4 w = make_float4(rr.x + e.alpha * v[j], rr.y + e.alpha * v[j + 1], rr.z + e.alpha * v[j + 2],
                               rr.w + e.alpha * v[j + 3]);
        *reinterpret_cast<float4*>(o + j) = w;
      }
    } else {
      for (int j = 0; j < 32 && n0 + j < N; ++j) o[j] = (r ? r[j] : 0.f) + e.alpha * v[j];
    }
  } else if (MODE == B2C_EPI_GLU_MASK) {
    // columns interleaved (value, gate). g = bf16(acc+bias) → out0 [M, N];
    // u[m, n/2] = bf16(a * sigmoid(b)) * (t < len) → out1 [M, N/2]      (conformer_modules.py:322-331)
    bf16* og = reinterpret_cast<bf16*>(e.out0) + (size_t)m * ldo + n0;
    bf16* ou = reinterpret_cast<bf16*>(e.out1) + (size_t)m * (ldo >> 1) + (n0 >> 1);
    bool valid = true;
    if (e.lengths) {
      int b = m / e.rows_per_batch, t = m - b * e.rows_per_batch;
      valid = t < e.lengths[b];
    }
    float u[16];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float a = bf16_round(v[j]), g = bf16_round(v[j + 1]);
      v[j] = a; v[j + 1] = g;
      u[j >> 1] = valid ? a * sigmoid_f(g) : 0.f;
    }
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        *reinterpret_cast<uint4*>(og + j) = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                                                       pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
#pragma unroll
      for (int j = 0; j < 16; j += 8)
        *reinterpret_cast<uint4*>(ou + j) = make_uint4(pack_bf16(u[j], u[j + 1]), pack_bf16(u[j + 2], u[j + 3]),
                                                       pack_bf16(u[j + 4], u[j + 5]), pack_bf16(u[j + 6], u[j + 7]));
    } else {
      for (int j = 0; j < 32 && n0 + j < N; ++j) og[j] = __float2bfloat16_rn(v[j]);
      for (int j = 0; j < 16 && n0 + 2 * j < N; ++j) ou[j] = __float2bfloat16_rn(u[j]);
    }
  } else if (MODE == B2C_EPI_DSILU) {
    // dz = bf16(acc * silu'(z[m,n]) [* dropout(site)])   z = aux (bf16)       (FFN backward)
    bf16* o = reinterpret_cast<bf16*>(e.out0) + (size_t)m * ldo + n0;
    const bf16* z = reinterpret_cast<const bf16*>(e.aux) + (size_t)m * ldo + n0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (n0 + j < N) {
        float g = bf16_round(v[j]);
        if (e.drop_keep != 0) {
          uint64_t idx = (uint64_t)m * (uint64_t)N + (uint64_t)(n0 + j);
          g = dropout_keep(e.seed, e.drop_site, idx, e.drop_keep) ? g * e.drop_scale : 0.f;
        }
        v[j] = g * dsilu_f(__bfloat162float(z[j]));
      }
    }
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        *reinterpret_cast<uint4*>(o + j) = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                                                      pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
    } else {
      for (int j = 0; j < 32 && n0 + j < N; ++j) o[j] = __float2bfloat16_rn(v[j]);
    }
  } else if (MODE == B2C_EPI_RELU_MASK) {
    // out_bf16 = relu(acc+bias) * (t < len)    rows are (b, t) x rows_inner  (subsampling.py:725-753)
    bf16* o = reinterpret_cast<bf16*>(e.out0) + (size_t)m * ldo + n0;
    bool valid = true;
    if (e.lengths) {
      int b = m / e.rows_per_batch, t = (m - b * e.rows_per_batch) / e.rows_inner;
      valid = t < e.lengths[b];
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = valid ? fmaxf(bf16_round(v[j]), 0.f) : 0.f;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        *reinterpret_cast<uint4*>(o + j) = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                                                      pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
    } else {
      for (int j = 0; j < 32 && n0 + j < N; ++j) o[j] = __float2bfloat16_rn(v[j]);
    }
  } else if (MODE == B2C_EPI_ACCUM_F32) {
    // out_f32[m,n] += alpha * acc  (weight gradients; atomic when split-K)
    float* o = reinterpret_cast<float*>(e.out0) + (size_t)m * ldo + n0;
    if (atomic) {
      for (int j = 0; j < 32 && n0 + j < N; ++j) atomicAdd(o + j, e.alpha * v[j]);
    } else if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 c = *reinterpret_cast<float4*>(o + j);
        c.x += e.alpha * v[j]; c.y += e.alpha * v[j + 1]; c.z += e.alpha * v[j + 2]; c.w += e.alpha * v[j + 3];
        *reinterpret_cast<float4*>(o + j) = c;
      }
    } else {
      for (int j = 0; j < 32 && n0 + j < N; ++j) o[j] += e.alpha * v[j];
    }
  }
}

// ---------------------------------------------------------------- kernel
template <int BN, int A_MN, int B_MN, int MODE>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using L = SmemLayout<BN>;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles * p.splits;
  const int kb_total = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile % p.splits;
        const int mn = tile / p.splits;
        const int m0 = (mn / n_tiles) * BM, n0 = (mn % n_tiles) * BN;
        const int kb0 = split * p.k_per_split;
        const int kb1 = min(kb_total, kb0 + p.k_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          mbar_expect_tx(&full_bar[stage], L::kStageBytes);
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * (64 * BK * 2), &tmA, &full_bar[stage], m0 + c * 64, kb * BK);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * (64 * BK * 2), &tmB, &full_bar[stage], n0 + c * 64, kb * BK);
          } else {
            tma_load_2d(sb, &tmB, &full_bar[stage], kb * BK, n0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0, phase = 0, as = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile % p.splits;
        const int kb0 = split * p.k_per_split;
        const int kb1 = min(kb_total, kb0 + p.k_per_split);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: +32 B per K=16 inside the 128 B swizzle row; SBO = 8 rows * 128 B.
            // MN-major: 16 k-rows * 128 B = 2048 B per K=16; LBO = stride between 64-wide MN chunks (64 k-rows*128 B).
            const uint64_t da = A_MN ? umma_desc_sw128(sa + k * 2048, 64 * BK * 2, 1024) : umma_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_desc_sw128(sb + k * 2048, 64 * BK * 2, 1024) : umma_desc_sw128(sb + k * 32, 16, 1024);
            umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete → epilogue
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int as = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mn = tile / p.splits;
      const int m0 = (mn / n_tiles) * BM, n0 = (mn % n_tiles) * BN;
      const int kb0 = (tile % p.splits) * p.k_per_split;
      const bool has_k = kb0 < kb_total;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const int m = m0 + q * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32), acc);
        tmem_ld_wait();
        if (has_k) epilogue_row<MODE>(p.epi, m, n0 + c * 32, p.M, p.N, acc, p.splits > 1);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

// ---------------------------------------------------------------- debug / tiny-shape SIMT GEMM (same epilogues)
// One thread per (row, 32-column chunk); used for shapes the TMA path cannot take (row pitch not 16 B aligned)
// and by the unit tests to bisect the tcgen05 path (B2C_GEMM_IMPL=simt).
template <int MODE>
__global__ void gemm_simt_kernel(const bf16* __restrict__ A, const bf16* __restrict__ B, int lda, int ldb, int a_mn,
                                 int b_mn, const GemmParams p) {
  const int chunks = (p.N + 31) / 32;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)p.M * chunks) return;
  const int m = (int)(gid / chunks), n0 = (int)(gid % chunks) * 32;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  for (int k = 0; k < p.K; ++k) {
    const float a = __bfloat162float(a_mn ? A[(size_t)k * lda + m] : A[(size_t)m * lda + k]);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int n = n0 + j;
      if (n < p.N) acc[j] += a * __bfloat162float(b_mn ? B[(size_t)k * ldb + n] : B[(size_t)n * ldb + k]);
    }
  }
  uint32_t r[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(acc[j]);
  epilogue_row<MODE>(p.epi, m, n0, p.M, p.N, r, false);
}

// ---------------------------------------------------------------- host launch
static int g_gemm_impl = -1;  // 0 = tcgen05, 1 = simt (debug)

template <int BN, int A_MN, int B_MN, int MODE>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using L = SmemLayout<BN>;
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    B2C_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles * p.splits;
  const int grid = total < num_sms() ? total : num_sms();
  kern<<<grid, kNumThreads, L::kTotal, stream>>>(ta, tb, p);
  B2C_LAUNCH_CHECK();
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
    const long long work = (long long)p.M * ((p.N + 31) / 32);
    gemm_simt_kernel<MODE><<<(unsigned)((work + 127) / 128), 128, 0, stream>>>(
        reinterpret_cast<const bf16*>(A), reinterpret_cast<const bf16*>(B), lda, ldb, a_mn, b_mn, p);
    B2C_LAUNCH_CHECK();
    return B2C_OK;
  }
  // tile width: 256 when it divides the work well, else 128
  const int BN = (p.N >= 256 && p.N % 256 == 0) ? 256 : 128;
  const int kb_total = (p.K + BK - 1) / BK;
  if (MODE == B2C_EPI_ACCUM_F32) {
    const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
    int splits = 1;
    if (tiles < num_sms()) splits = (num_sms() + tiles - 1) / tiles;
    if (splits > kb_total) splits = kb_total;
    if (splits < 1) splits = 1;
    p.k_per_split = (kb_total + splits - 1) / splits;
    p.splits = (kb_total + p.k_per_split - 1) / p.k_per_split;
  } else {
    p.splits = 1;
    p.k_per_split = kb_total;
  }
  CUtensorMap ta, tb;
  {
    // K-major: dims {K, rows}, box {64, 128|BN};  MN-major: dims {rows, K}, box {64, 64}
    uint64_t dims[2], strides[1];
    uint32_t box[2], es[2] = {1, 1};
    if (a_mn) { dims[0] = p.M; dims[1] = p.K; box[0] = 64; box[1] = BK; }
    else      { dims[0] = p.K; dims[1] = p.M; box[0] = BK; box[1] = BM; }
    strides[0] = (uint64_t)lda * 2;
    int rc = make_tmap(&ta, A, 2, 2, dims, strides, box, es, 1);
    if (rc) return rc;
    if (b_mn) { dims[0] = p.N; dims[1] = p.K; box[0] = 64; box[1] = BK; }
    else      { dims[0] = p.K; dims[1] = p.N; box[0] = BK; box[1] = BN; }
    strides[0] = (uint64_t)ldb * 2;
    rc = make_tmap(&tb, B, 2, 2, dims, strides, box, es, 1);
    if (rc) return rc;
  }
#define B2C_GO(BN_, AM_, BM_) return launch_tc<BN_, AM_, BM_, MODE>(ta, tb, p, stream)
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
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.splits = 1; p.k_per_split = 0; p.epi = *epi;
  switch (mode) {
    case B2C_EPI_BF16: return dispatch_mode<B2C_EPI_BF16>(A, B, lda, ldb, a_mn, b_mn, p, stream);
    case B2C_EPI_SILU: return dispatch_mode<B2C_EPI_SILU>(A, B, lda, ldb, a_mn, b_mn, p, stream);
    case B2C_EPI_RESID_F32: return dispatch_mode<B2C_EPI_RESID_F32>(A, B, lda, ldb, a_mn, b_mn, p, stream);
    case B2C_EPI_GLU_MASK: return dispatch_mode<B2C_EPI_GLU_MASK>(A, B, lda, ldb, a_mn, b_mn, p, stream);
    case B2C_EPI_DSILU: return dispatch_mode<B2C_EPI_DSILU>(A, B, lda, ldb, a_mn, b_mn, p, stream);
    case B2C_EPI_RELU_MASK: return dispatch_mode<B2C_EPI_RELU_MASK>(A, B, lda, ldb, a_mn, b_mn, p, stream);
    case B2C_EPI_ACCUM_F32: return dispatch_mode<B2C_EPI_ACCUM_F32>(A, B, lda, ldb, a_mn, b_mn, p, stream);
  }
  set_last_error("gemm: unknown epilogue %d", mode);
  return B2C_ERR_INVALID;
}

}  // namespace b2c

extern "C" int b2c_gemm(const void* A, const void* B, int M, int N, int K, int lda, int ldb, int a_mn_major,
                        int b_mn_major, int epilogue_mode, const b2c_epilogue* epi, void* stream) {
  return b2c::gemm(A, B, M, N, K, lda, ldb, a_mn_major, b_mn_major, epilogue_mode, epi, (cudaStream_t)stream);
}
