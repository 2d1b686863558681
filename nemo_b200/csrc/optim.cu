// The step on the other side of the gradient all-reduce and the per-step weight preparation, both as
// table-driven multi-tensor kernels: ONE launch walks every parameter tensor of the encoder.
//
//   b2c_adamw_*      AdamW (decoupled weight decay; torch.optim.AdamW arithmetic, the optimizer NeMo builds for
//                    `optim.name: adamw`, nemo/core/classes/modelPT.py:650-823, examples/asr/conf/conformer/
//                    conformer_ctc_bpe.yaml:175-192: lr 2.0, betas 0.9/0.98, weight_decay 1e-3) with the NoamAnnealing
//                    learning-rate policy (nemo/core/optim/lr_scheduler.py:518-575) evaluated ON THE DEVICE from a
//                    device-resident step counter, so the whole training step can be captured in a CUDA graph.
//                    Consumes the flat fp32 gradient buffer the backward kernels wrote (nemo_b200/parallel.py).
//   b2c_prep_weights fp32 master weights → bf16 (or fp32) operand copies in the layouts the GEMM / conv kernels read:
//                    plain casts and 3-D index permutations (GLU value/gate interleave, conv taps-major, flatten order)
//                    described by per-tensor output strides.
//   b2c_counter_add  advances the device-side RNG counter (common.cuh: seed_eff) — first node of a captured step.
#include "common.cuh"

namespace b2c {

int set_seed_counter_gemm(const unsigned long long* p);
int set_seed_counter_attn(const unsigned long long* p);
int set_seed_counter_attn_tc(const unsigned long long* p);
int set_seed_counter_elementwise(const unsigned long long* p);
int set_seed_counter_mel(const unsigned long long* p);

constexpr int kChunk = 2048;  // elements per block (256 threads x 8)

__global__ void counter_add_kernel(unsigned long long* ctr, unsigned long long inc) { *ctr += inc; }

// ---------------------------------------------------------------- AdamW + Noam
// state[0] = optimizer step k (as double, 1-based after the increment); scal[0..3] = lr, lr / bias_correction1,
// 1 / sqrt(bias_correction2), 1 - lr * weight_decay
__global__ void adamw_prepare_kernel(double* __restrict__ state, float* __restrict__ scal, b2c_adamw_hyper h) {
  const double k = state[0] + 1.0;
  state[0] = k;
  double lr = (double)h.lr;
  if (h.noam_d_model > 0) {
    // torch's LRScheduler computes the rate for optimizer step k from last_epoch = k - 1; NoamAnnealing.get_lr uses
    // step = max(1, last_epoch) (lr_scheduler.py:555), so steps 1 and 2 share step = 1
    const double step = k - 1.0 > 1.0 ? k - 1.0 : 1.0;
    const double norm = pow((double)h.noam_d_model, -0.5);
    double mult;
    if (h.noam_warmup > 0) {
      const double a = pow(step, -0.5), b = step * pow((double)h.noam_warmup, -1.5);
      mult = norm * (a < b ? a : b);
    } else {
      mult = norm * pow(step, -0.5);
    }
    lr = (double)h.lr * mult;
    if (step > (double)h.noam_warmup && lr < (double)h.noam_min_lr) lr = (double)h.noam_min_lr;  // lr_scheduler.py:573-574
  }
  const double bc1 = 1.0 - pow((double)h.beta1, k), bc2 = 1.0 - pow((double)h.beta2, k);
  scal[0] = (float)lr;
  scal[1] = (float)(lr / bc1);
  scal[2] = (float)(1.0 / sqrt(bc2));
  scal[3] = (float)(1.0 - lr * (double)h.weight_decay);
}

__device__ __forceinline__ void adamw_one(float& p, float& m, float& v, float g, float b1, float b2, float eps, float step_size,
                                          float inv_sqrt_bc2, float decay) {
  // torch/optim/adamw.py (_single_tensor_adamw): p *= 1 - lr*wd; m = lerp(m, g, 1-b1); v = b2*v + (1-b2) g^2;
  // p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
  p *= decay;
  m = m + (g - m) * (1.f - b1);
  v = b2 * v + (1.f - b2) * g * g;
  const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256)
adamw_apply_kernel(const b2c_tensor_entry* __restrict__ table, const int2* __restrict__ blocks, float* __restrict__ grad,
                   float* __restrict__ m_, float* __restrict__ v_, const float* __restrict__ scal, float beta1, float beta2,
                   float eps, float grad_scale, int zero_grad) {
  const int2 bk = blocks[blockIdx.x];
  const b2c_tensor_entry e = table[bk.x];
  const long long i0 = (long long)bk.y * kChunk;
  const int n = (int)min((long long)kChunk, e.numel - i0);
  float* p = reinterpret_cast<float*>(e.ptr) + i0;
  float* g = grad + e.offset + i0;
  float* m = m_ + e.offset + i0;
  float* v = v_ + e.offset + i0;
  const float step_size = scal[1], inv_sqrt_bc2 = scal[2], decay = e.flags & 1 ? 1.f : scal[3];
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
  if (vec) {
    for (int i = threadIdx.x * 4; i < (n & ~3); i += 256 * 4) {
      float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      adamw_one(pp.x, mm.x, vv.x, gg.x * grad_scale, beta1, beta2, eps, step_size, inv_sqrt_bc2, decay);
      adamw_one(pp.y, mm.y, vv.y, gg.y * grad_scale, beta1, beta2, eps, step_size, inv_sqrt_bc2, decay);
      adamw_one(pp.z, mm.z, vv.z, gg.z * grad_scale, beta1, beta2, eps, step_size, inv_sqrt_bc2, decay);
      adamw_one(pp.w, mm.w, vv.w, gg.w * grad_scale, beta1, beta2, eps, step_size, inv_sqrt_bc2, decay);
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
      if (zero_grad) *reinterpret_cast<float4*>(g + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = (n & ~3) + threadIdx.x; i < n; i += 256) {
      adamw_one(p[i], m[i], v[i], g[i] * grad_scale, beta1, beta2, eps, step_size, inv_sqrt_bc2, decay);
      if (zero_grad) g[i] = 0.f;
    }
  } else {
    for (int i = threadIdx.x; i < n; i += 256) {
      adamw_one(p[i], m[i], v[i], g[i] * grad_scale, beta1, beta2, eps, step_size, inv_sqrt_bc2, decay);
      if (zero_grad) g[i] = 0.f;
    }
  }
}

// ---------------------------------------------------------------- weight preparation
// entry: src fp32 [A, B, C] (contiguous), dst element (a, b, c) at a*sa + b*sb + c*sc; dst is bf16 unless flags & 2 (fp32)
__global__ void __launch_bounds__(256)
prep_weights_kernel(const b2c_prep_entry* __restrict__ table, const int2* __restrict__ blocks) {
  const int2 bk = blocks[blockIdx.x];
  const b2c_prep_entry e = table[bk.x];
  const long long i0 = (long long)bk.y * kChunk;
  const long long numel = (long long)e.A * e.B * e.C;
  const int n = (int)min((long long)kChunk, numel - i0);
  const float* src = reinterpret_cast<const float*>(e.src);
  const bool f32 = (e.flags & 2) != 0;
  const bool plain = (e.sc == 1 && e.sb == e.C && e.sa == (long long)e.B * e.C);
  if (plain && !f32 && ((((uintptr_t)(src + i0)) & 15) == 0) && (((uintptr_t)(reinterpret_cast<bf16*>(e.dst) + i0)) & 7) == 0) {
    bf16* dst = reinterpret_cast<bf16*>(e.dst) + i0;
    for (int i = threadIdx.x * 4; i < (n & ~3); i += 256 * 4) {
      const float4 x = *reinterpret_cast<const float4*>(src + i0 + i);
      *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
    }
    for (int i = (n & ~3) + threadIdx.x; i < n; i += 256) dst[i] = __float2bfloat16_rn(src[i0 + i]);
    return;
  }
  const long long BC = (long long)e.B * e.C;
  for (int i = threadIdx.x; i < n; i += 256) {
    const long long idx = i0 + i;
    const long long a = idx / BC, r = idx - a * BC;
    const long long b = r / e.C, c = r - b * e.C;
    const long long o = a * e.sa + b * e.sb + c * e.sc;
    const float x = src[idx];
    if (f32) reinterpret_cast<float*>(e.dst)[o] = x;
    else reinterpret_cast<bf16*>(e.dst)[o] = __float2bfloat16_rn(x);
  }
}

}  // namespace b2c

using namespace b2c;

extern "C" int b2c_counter_add(unsigned long long* counter, unsigned long long inc, void* stream) {
  B2C_REQUIRE(counter != nullptr, "counter_add: null counter");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, inc);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_set_seed_counter(const unsigned long long* counter) {
  int rc;
  if ((rc = set_seed_counter_gemm(counter))) return rc;
  if ((rc = set_seed_counter_attn(counter))) return rc;
  if ((rc = set_seed_counter_attn_tc(counter))) return rc;
  if ((rc = set_seed_counter_elementwise(counter))) return rc;
  if ((rc = set_seed_counter_mel(counter))) return rc;
  return B2C_OK;
}

extern "C" int b2c_adamw_step(const b2c_tensor_entry* table, const int32_t* blocks, int n_blocks, float* grad_flat,
                              float* exp_avg, float* exp_avg_sq, double* state, float* scalars, const b2c_adamw_hyper* hyper,
                              float grad_scale, int zero_grad, void* stream) {
  B2C_REQUIRE(table && blocks && grad_flat && exp_avg && exp_avg_sq && state && scalars && hyper, "adamw_step: null argument");
  B2C_REQUIRE(n_blocks > 0, "adamw_step: empty block table");
  B2C_REQUIRE(hyper->beta1 >= 0.f && hyper->beta1 < 1.f && hyper->beta2 >= 0.f && hyper->beta2 < 1.f && hyper->eps > 0.f,
              "adamw_step: bad hyper-parameters");
  adamw_prepare_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, scalars, *hyper);
  B2C_LAUNCH_CHECK();
  adamw_apply_kernel<<<n_blocks, 256, 0, (cudaStream_t)stream>>>(table, reinterpret_cast<const int2*>(blocks), grad_flat, exp_avg,
                                                                 exp_avg_sq, scalars, hyper->beta1, hyper->beta2, hyper->eps,
                                                                 grad_scale, zero_grad);
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}

extern "C" int b2c_prep_weights(const b2c_prep_entry* table, const int32_t* blocks, int n_blocks, void* stream) {
  B2C_REQUIRE(table && blocks && n_blocks > 0, "prep_weights: empty table");
  prep_weights_kernel<<<n_blocks, 256, 0, (cudaStream_t)stream>>>(table, reinterpret_cast<const int2*>(blocks));
  B2C_LAUNCH_CHECK();
  return B2C_OK;
}
