// Fused optimizer step over flat fp32 buffers (runner.py:411-427: grad /= n, global-norm clip,
// Adam, zero_grad) -- one HBM pass: reads p, g, m, v; writes p, m, v (and zeros g).
#include "mh_b200.h"
#include "mh_common.cuh"

namespace mh {
extern long long g_launches;

// keep-mask of 4 consecutive elements (one byte each, 0 = pruned) -> the 4 values with pruned ones zeroed
__device__ __forceinline__ float4 mask4(float4 v, uint32_t m) {
  if ((m & 0xFFu) == 0) v.x = 0.f;
  if ((m & 0xFF00u) == 0) v.y = 0.f;
  if ((m & 0xFF0000u) == 0) v.z = 0.f;
  if ((m & 0xFF000000u) == 0) v.w = 0.f;
  return v;
}

// mask (optional, one byte per element of the flat buffer): gradients of pruned elements do not count (the reference
// zeroes them through the masked_fill of pytorch_code/prune.py:38 in the autograd graph)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask, long long n,
                                                    float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (; i + 4 <= n; i += stride) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x + i));
    if (mask != nullptr) v = mask4(v, __ldg(reinterpret_cast<const uint32_t*>(mask + i)));
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (i < n && i + 4 > n)
    for (long long j = i; j < n; ++j) s += (mask == nullptr || mask[j]) ? x[j] * x[j] : 0.f;
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = red[threadIdx.x];
    t += __shfl_xor_sync(0xffu, t, 4);
    t += __shfl_xor_sync(0xffu, t, 2);
    t += __shfl_xor_sync(0xffu, t, 1);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
            float lr, float b1, float b2, float eps, float wd, const unsigned long long* __restrict__ step_ptr, float grad_scale,
            float max_norm, const float* __restrict__ sumsq, int zero_grad, __nv_bfloat16* __restrict__ shadow,
            const uint8_t* __restrict__ mask, float* __restrict__ eff) {
  const float step = static_cast<float>(*step_ptr);
  float clip = 1.f;
  bool skip = false;
  if (sumsq != nullptr) {
    const float norm = sqrtf(*sumsq) * grad_scale;
    if (!(norm == norm) || isinf(norm)) skip = true;  // runner.py:417-425: NaN grad norm skips the step
    if (max_norm > 0.f) clip = fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const float gs = grad_scale * clip;
  const float bc1 = 1.f - powf(b1, step);
  const float bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (; i < n; i += stride) {  // n is padded to a multiple of 4 by the host
    float4 pv = *reinterpret_cast<float4*>(p + i), gv = *reinterpret_cast<float4*>(g + i);
    float4 mv = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
    const uint32_t mk = mask != nullptr ? __ldg(reinterpret_cast<const uint32_t*>(mask + i)) : 0x01010101u;
    if (mask != nullptr) gv = mask4(gv, mk);  // pruned elements receive no gradient (their moments keep decaying)
    if (!skip) {
      float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float gj = gp[j] * gs;
        if (wd != 0.f) gj += wd * pp[j];
        mp[j] = b1 * mp[j] + (1.f - b1) * gj;
        vp[j] = b2 * vp[j] + (1.f - b2) * gj * gj;
        const float denom = sqrtf(vp[j]) * inv_sqrt_bc2 + eps;
        pp[j] -= step_size * mp[j] / denom;
      }
      *reinterpret_cast<float4*>(p + i) = pv;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(v + i) = vv;
      // bf16 shadow of the updated parameters = next step's GEMM operands (replaces the per-step weight prep)
      // (with a prune mask: the EFFECTIVE parameters weight_orig * mask of pytorch_code/prune.py:38, which is what the
      // 144 per-forward masked_fill launches of the reference recompute; `eff` is their fp32 copy for the bias operands)
      const float4 ev = mask != nullptr ? mask4(pv, mk) : pv;
      if (shadow != nullptr)
        *reinterpret_cast<uint2*>(shadow + i) = make_uint2(f32x2_to_bf16(ev.x, ev.y), f32x2_to_bf16(ev.z, ev.w));
      if (eff != nullptr) *reinterpret_cast<float4*>(eff + i) = ev;
    }
    if (zero_grad) *reinterpret_cast<float4*>(g + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
}  // namespace mh

using namespace mh;
#define ST reinterpret_cast<cudaStream_t>(stream)

static int sumsq_impl(const float* x, const uint8_t* mask, long long n, float* out, void* stream);
extern "C" int mh_sumsq(const float* x, long long n, float* out, void* stream) { return sumsq_impl(x, nullptr, n, out, stream); }
extern "C" int mh_sumsq_masked(const float* x, const uint8_t* mask, long long n, float* out, void* stream) {
  MH_CHECK(mask == nullptr || (reinterpret_cast<uintptr_t>(mask) & 3) == 0, "sumsq: the mask must be 4-byte aligned");
  return sumsq_impl(x, mask, n, out, stream);
}
static int sumsq_impl(const float* x, const uint8_t* mask, long long n, float* out, void* stream) {
  if (n == 0) return 0;
  long long g = (n / 4 + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  sumsq_kernel<<<static_cast<int>(g), 256, 0, ST>>>(x, mask, n, out);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

static int adam_impl(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, const unsigned long long* step, float grad_scale, float max_norm,
                     const float* sumsq, int zero_grad, void* bf16_shadow, const uint8_t* mask, float* eff, void* stream) {
  MH_CHECK(n % 4 == 0, "adam: flat buffer length must be a multiple of 4 (got %lld)", n);
  MH_CHECK(mask == nullptr || (reinterpret_cast<uintptr_t>(mask) & 3) == 0, "adam: the mask must be 4-byte aligned");
  if (n == 0) return 0;
  long long g = (n / 4 + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (g > cap) g = cap;
  adam_kernel<<<static_cast<int>(g), 256, 0, ST>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                   step, grad_scale, max_norm, sumsq, zero_grad,
                                                   reinterpret_cast<__nv_bfloat16*>(bf16_shadow), mask, eff);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

extern "C" int mh_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                            float beta1, float beta2, float eps, float weight_decay, const unsigned long long* step, float grad_scale,
                            float max_norm, const float* sumsq, int zero_grad, void* bf16_shadow, void* stream) {
  return adam_impl(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, max_norm, sumsq,
                   zero_grad, bf16_shadow, nullptr, nullptr, stream);
}

extern "C" int mh_adam_step_masked(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, const unsigned long long* step,
                                   float grad_scale, float max_norm, const float* sumsq, int zero_grad, void* bf16_shadow,
                                   const uint8_t* mask, float* effective, void* stream) {
  return adam_impl(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, max_norm, sumsq,
                   zero_grad, bf16_shadow, mask, effective, stream);
}

// shadow / effective copies of a flat parameter buffer outside the optimizer (after load_state_dict, prune events, ...)
__global__ void __launch_bounds__(256)
flat_effective_kernel(const float* __restrict__ p, const uint8_t* __restrict__ mask, __nv_bfloat16* __restrict__ shadow,
                      float* __restrict__ eff, long long n) {
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 4;
  for (; i < n; i += stride) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p + i));
    if (mask != nullptr) v = mh::mask4(v, __ldg(reinterpret_cast<const uint32_t*>(mask + i)));
    if (shadow != nullptr) *reinterpret_cast<uint2*>(shadow + i) = make_uint2(f32x2_to_bf16(v.x, v.y), f32x2_to_bf16(v.z, v.w));
    if (eff != nullptr) *reinterpret_cast<float4*>(eff + i) = v;
  }
}

extern "C" int mh_flat_effective(const float* param, const uint8_t* mask, void* bf16_shadow, float* effective, long long n,
                                 void* stream) {
  MH_CHECK(n % 4 == 0 && (mask == nullptr || (reinterpret_cast<uintptr_t>(mask) & 3) == 0), "flat_effective: n %% 4 / mask alignment");
  if (n == 0) return 0;
  long long g = (n / 4 + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (g > cap) g = cap;
  flat_effective_kernel<<<static_cast<int>(g), 256, 0, ST>>>(param, mask, reinterpret_cast<__nv_bfloat16*>(bf16_shadow), effective, n);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}
