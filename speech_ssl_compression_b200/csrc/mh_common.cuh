// Shared helpers for all kernels: error plumbing for the C ABI, the counter-based dropout
// RNG (Philox4x32-7), fp32 erf-GELU, small vector load/store helpers.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace mh {

// ---- error plumbing (thread-local message, returned through mh_last_error()) ----
void set_error(const char* fmt, ...);
#define MH_CHECK(cond, ...)            \
  do {                                 \
    if (!(cond)) {                     \
      ::mh::set_error(__VA_ARGS__);    \
      return 1;                        \
    }                                  \
  } while (0)
#define MH_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::mh::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return static_cast<int>(_e);                                                            \
    }                                                                                         \
  } while (0)
#define MH_LAUNCH_CHECK() MH_CUDA(cudaGetLastError())

int sm_count();

// ---- programmatic dependent launch (PDL) ----
// A step is ~320 short kernels and ~1.0 of its 21.4 ms is launch gap.  Kernels launched through launch_pdl() may be
// scheduled while the previous kernel of the stream is still draining; every such kernel executes
// griddepcontrol.wait (all prerequisite grids complete, their writes visible) before it touches global memory, so
// launch latency, CTA scheduling and the set-up in front of the wait can overlap the predecessor's tail.
// MEASURED (bench.py, 1 x B200, CUDA graph): neutral for the GEMM and attention kernels (21.40 vs 21.34 ms per step,
// within box-to-box noise), 0.56 ms SLOWER for the many-small-CTA norm kernels -- the graph already hides most of the
// launch latency.  Hence off by default; MH_PDL = bit mask of kernel families (1 GEMM, 2 attention, 4 norm).
bool pdl_enabled(int family = 1);
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_launch_dependents();
  pdl_wait();
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_f(int family, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled(family) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define launch_pdl(...) launch_pdl_f(MH_PDL_FAMILY, __VA_ARGS__)

// ---- Philox4x32-7 (Salmon et al., SC'11: 7 rounds is the smallest crush-resistant variant) ----
// Dropout over a logical element stream.  One Philox call covers 8 consecutive elements (eight 16-bit lanes, low
// half of each word first); element e is kept iff lane(e) >= thresh, thresh = round(p * 65536).  The forward and
// the backward regenerate identical decisions from (seed, device step counter, site, element / 8):
//   key     = seed                      -> the 7 round-key pairs are computed ON THE HOST (make_drop) and travel in
//                                          the kernel parameters: they are constant-bank operands of the LOP3s, no
//                                          registers and no per-call key schedule
//   counter = (group_lo, group_hi ^ step_lo, site, 0x9E3779B9 ^ step_hi); `step` is the optional device counter
//             (CUDA-graph replays advance it), read once per thread (DropState).
// A call is 14 IMAD.WIDE + 14 LOP3 (+1) for 8 decisions.  (The first version folded the device counter into the
// KEY: every call re-derived the key schedule and re-read the counter, ~100 instructions per 8 elements, 45 % of
// the instruction-bound GELU / dropout GEMM epilogues and a third of the LayerNorm backward.)
struct DropCfg {
  uint32_t ka[7], kb[7];             // Philox round keys: ka[r] = seed_lo + r * 0x9E3779B9, kb[r] = seed_hi + r * 0xBB67AE85
  const unsigned long long* offset;  // optional device step counter (CUDA-graph replays)
  uint32_t site;    // unique per dropout site (layer * 8 + site id)
  uint32_t thresh;  // 0 => dropout disabled
  float scale;      // 1 / (1 - p)
};
const unsigned long long* dropout_offset_ptr();
__host__ inline DropCfg make_drop(float p, uint64_t seed, uint32_t site) {
  DropCfg d;
  uint32_t a = static_cast<uint32_t>(seed), b = static_cast<uint32_t>(seed >> 32);
  for (int r = 0; r < 7; ++r) { d.ka[r] = a; d.kb[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
  d.offset = dropout_offset_ptr();
  d.site = site;
  d.thresh = p > 0.f ? static_cast<uint32_t>(p * 65536.0f + 0.5f) : 0u;
  if (d.thresh > 65535u) d.thresh = 65535u;
  d.scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  return d;
}

// Per-thread part of a dropout stream: the device step counter, read once (2 registers).  Every call takes the
// DropCfg it was built from -- pass the kernel parameter itself so that keys / site / threshold stay constant-bank
// operands.
struct DropState {
  uint32_t c1, c3;
  __device__ __forceinline__ explicit DropState(const DropCfg& d) {
    unsigned long long o = 0ull;
    if (d.thresh != 0 && d.offset != nullptr) o = __ldg(d.offset);
    c1 = static_cast<uint32_t>(o);
    c3 = 0x9E3779B9u ^ static_cast<uint32_t>(o >> 32);
  }
  // the 128 random bits of `group` (lane j of the group = 16-bit field j, low half first)
  __device__ __forceinline__ uint4 bits(const DropCfg& d, uint64_t group) const {
    uint32_t x0 = static_cast<uint32_t>(group), x1 = static_cast<uint32_t>(group >> 32) ^ c1, x2 = d.site, x3 = c3;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * x0;
      const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * x2;
      const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ x1 ^ d.ka[r], n2 = static_cast<uint32_t>(p0 >> 32) ^ x3 ^ d.kb[r];
      x1 = static_cast<uint32_t>(p1); x3 = static_cast<uint32_t>(p0);
      x0 = n0; x2 = n2;
    }
    return make_uint4(x0, x1, x2, x3);
  }
  // v[j] = keep_j ? v[j] * scale : 0 for the 8 elements of `group`, without materialising the bit mask:
  // the high field is compared in place (word >= thresh << 16), the low one after a 16-bit shift.
  __device__ __forceinline__ void apply8(const DropCfg& d, uint64_t group, float (&v)[8]) const {
    const uint4 r = bits(d, group);
    const uint32_t t = d.thresh << 16;
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = (w[i] << 16) >= t ? v[2 * i] * d.scale : 0.f;
      v[2 * i + 1] = w[i] >= t ? v[2 * i + 1] * d.scale : 0.f;
    }
  }
  // keep-mask bits (bit j = element j kept) of the 8 elements of `group`
  __device__ __forceinline__ uint32_t keep8(const DropCfg& d, uint64_t group) const {
    const uint4 r = bits(d, group);
    const uint32_t t = d.thresh << 16;
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      m |= ((w[i] << 16) >= t ? 1u : 0u) << (2 * i);
      m |= (w[i] >= t ? 1u : 0u) << (2 * i + 1);
    }
    return m;
  }
};

// ---- math ----
// erf-GELU in fp32 (fairseq_code/gelu.py:34-35) without erff():
//   erfc(a / sqrt 2) = 2^-g(a),  g(a) ~= a (c0 + c1 a + c2 a^2 + c3 a^3 + c4 a^4)  (least-squares fit on
//   [0, 9], weighted by the sensitivity of gelu; |gelu error| < 7e-7 absolute over the whole line, i.e.
//   below fp32 round-off of the reference's own erff for |x| > 1e-2 and 4 decimal orders below one bf16 ulp)
//   gelu(x) = max(x, 0) - 0.5 |x| erfc(|x| / sqrt 2);   Phi(x) = x >= 0 ? 1 - e/2 : e/2.
__device__ __forceinline__ float erfc_half_scaled(float a) {  // 0.5 * erfc(a / sqrt(2)), a >= 0
  // no clamp needed: the quartic grows monotonically past the fit range, so 2^-(...) underflows to 0
  float p = 4.88221852e-04f;
  p = fmaf(p, a, -7.19561887e-03f);
  p = fmaf(p, a, 5.21302448e-02f);
  p = fmaf(p, a, 4.59620056e-01f);
  p = fmaf(p, a, 1.15099005e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(-p, a, -1.0f)));  // 2^(-g - 1)
  return e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float a = fabsf(x);
  // |x| e(|x|) -> 0 for large finite |x| (e underflows long before |x| overflows); x = +-inf is not a valid activation
  return fmaf(-a, erfc_half_scaled(a), fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float h = erfc_half_scaled(fabsf(x));
  const float cdf = x >= 0.f ? 1.0f - h : h;
  float pdf;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pdf) : "f"(-0.72134752044f * x * x));
  return fmaf(x * 0.39894228040143268f, pdf, cdf);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- 128-bit helpers ----
__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg128(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ void bf16x8_to_f32(uint4 v, float (&f)[8]) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xFFFF0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xFFFF0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xFFFF0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xFFFF0000u);
}
__device__ __forceinline__ uint32_t f32x2_to_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint4 f32_to_bf16x8(const float (&f)[8]) {
  return make_uint4(f32x2_to_bf16(f[0], f[1]), f32x2_to_bf16(f[2], f[3]), f32x2_to_bf16(f[4], f[5]),
                    f32x2_to_bf16(f[6], f[7]));
}

}  // namespace mh
