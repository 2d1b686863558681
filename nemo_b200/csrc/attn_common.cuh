// Shared declarations of the rel-pos attention kernels (mma.sync and tcgen05 variants).
#pragma once
#include "common.cuh"

namespace b2c {

constexpr float kLog2e = 1.4426950408889634f;

struct AttnArgs {
  const bf16* qkv;   // [B*T, 3d]
  const bf16* pos;   // [2T-1, d]
  const float* bias_u;
  const float* bias_v;
  const int32_t* lengths;
  int B, T, H, d;
  float scale;       // 1/sqrt(dk)
  float drop_p;
  unsigned int drop_keep;
  float drop_scale;
  unsigned long long seed;
  unsigned int site;
};


// keep decisions for the element pair (i, j), (i, j+1) with j even: bit0 / bit1
__device__ __forceinline__ uint32_t attn_keep2(const AttnArgs& a, uint64_t row_base, int j) {
  const uint32_t hsh = dropout_hash(a.seed, a.site, (row_base + (uint64_t)j) >> 1);
  return ((hsh & 0xFFFFu) < a.drop_keep ? 1u : 0u) | ((hsh >> 16) < a.drop_keep ? 2u : 0u);
}
// rows are padded to a multiple of 16 keys so that a row's 16-column dropout chunks are chunks of the flat index
__device__ __forceinline__ uint64_t attn_row_base(const AttnArgs& a, int b, int h, int i) {
  return ((uint64_t)((b * a.H + h) * (int64_t)a.T + i)) * (uint64_t)((a.T + 15) & ~15);
}
// chunk hash of the 16 keys starting at j (j % 16 == 0); pair p of the chunk: dropout_word(base, p)
__device__ __forceinline__ uint32_t attn_keep_base(const AttnArgs& a, uint64_t row_base, int j) {
  return dropout_base(a.seed, a.site, (row_base + (uint64_t)j) >> 4);
}

// tcgen05 / TMEM / TMA forward (attention_tc.cu); dk == 64 only
int launch_relpos_attn_fwd_tc(const AttnArgs& a, bf16* ctx, float* ctx32, float* lse, cudaStream_t st);
int launch_relpos_attn_bwd_q_tc(const AttnArgs& a, const bf16* ctx, const float* ctx32, const bf16* dctx, const float* lse,
                                bf16* dqkv, float* dbu, float* dbv, bf16* ws_ds, bf16* ws_p, bf16* ws_qu, int Tp, cudaStream_t st);
int launch_relpos_attn_bwd_kv_tc(const AttnArgs& a, const bf16* dctx, const bf16* ws_ds, const bf16* ws_p, const bf16* ws_qu,
                                 bf16* dqkv, int Tp, cudaStream_t st);

}  // namespace b2c
