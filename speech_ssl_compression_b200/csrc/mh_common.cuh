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
// Bit-sliced dropout over a logical element stream.  The stream is cut into WORDS of 32 consecutive elements; one
// thread produces the 32 keep decisions of a word at once:
//   * every element gets a 12-bit random number R, stored bit-sliced: plane k is a 32-bit word whose bit i is bit k of
//     element i's number.  Planes 4..11 (the eight most significant bits) are the 8 output words of two Philox4x32-7
//     calls; planes 0..3 only matter to the 1 element in 256 whose upper bits equal the threshold's, and are lane
//     rotations of the upper planes -- for such an element its own upper bits are fixed, so its low bits are the
//     (independent, uniform) upper bits of the elements 11 and 19 positions away: unbiased at 3 instead of 7
//     instructions per plane;
//   * element i is kept iff R_i >= thr, thr = round(p * 4096): a 12-step LOP3 chain, one instruction per plane for all
//     32 decisions (tmask[k] = all-ones when bit k of thr is set); survivors are scaled by 4096 / (4096 - thr).
//   p is therefore quantised to 1/4096 (0.1 -> 410/4096 = 0.100098).
// 2 x (14 IMAD.WIDE + 14 LOP3) + 12 + 12 instructions for 32 decisions = 2.5 per element; the first scheme (one call
// per 8 elements, 16-bit lane compares) cost 8.5 per element and was 35 % of the instruction-bound GELU / dropout
// GEMM epilogues.  The forward and the backward regenerate identical decisions from (seed, device step counter,
// site, word index):
//   key     = seed                      -> the 7 round-key pairs are computed ON THE HOST (make_drop) and travel in
//                                          the kernel parameters: constant-bank operands of the LOP3s
//   counter = (c_lo, c_hi ^ step_lo, site, 0x9E3779B9 ^ step_hi), c = 2 * word + call; `step` is the optional device
//             counter (CUDA-graph replays advance it), read once per thread (DropState).
struct DropCfg {
  uint32_t ka[7], kb[7];             // Philox round keys: ka[r] = seed_lo + r * 0x9E3779B9, kb[r] = seed_hi + r * 0xBB67AE85
  const unsigned long long* offset;  // optional device step counter (CUDA-graph replays)
  uint32_t site;    // unique per dropout site (layer * 8 + site id)
  uint32_t thresh;  // thr = round(p * 4096); 0 => dropout disabled
  float scale;      // 4096 / (4096 - thr)
  uint32_t tmask[12];  // plane k: all-ones when bit k of thr is set
};
const unsigned long long* dropout_offset_ptr();
__host__ inline DropCfg make_drop(float p, uint64_t seed, uint32_t site) {
  DropCfg d;
  uint32_t a = static_cast<uint32_t>(seed), b = static_cast<uint32_t>(seed >> 32);
  for (int r = 0; r < 7; ++r) { d.ka[r] = a; d.kb[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
  d.offset = dropout_offset_ptr();
  d.site = site;
  d.thresh = p > 0.f ? static_cast<uint32_t>(p * 4096.0f + 0.5f) : 0u;
  if (p > 0.f && d.thresh == 0u) d.thresh = 1u;
  if (d.thresh > 4095u) d.thresh = 4095u;
  d.scale = 4096.0f / static_cast<float>(4096u - d.thresh);
  for (int k = 0; k < 12; ++k) d.tmask[k] = ((d.thresh >> k) & 1u) ? 0xffffffffu : 0u;
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
  // one Philox4x32-7 block
  __device__ __forceinline__ uint4 bits(const DropCfg& d, uint64_t ctr) const {
    uint32_t x0 = static_cast<uint32_t>(ctr), x1 = static_cast<uint32_t>(ctr >> 32) ^ c1, x2 = d.site, x3 = c3;
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
  // the 32 keep decisions of stream word `word`: bit i = element 32 word + i is kept
  __device__ __forceinline__ uint32_t keep32(const DropCfg& d, uint64_t word) const {
    uint32_t w[12];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const uint4 b = bits(d, 2 * word + t);
      w[4 + 4 * t] = b.x; w[5 + 4 * t] = b.y; w[6 + 4 * t] = b.z; w[7 + 4 * t] = b.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = __funnelshift_l(w[4 + k], w[4 + k], 11) ^ __funnelshift_l(w[8 + k], w[8 + k], 19);
    uint32_t ge = 0xffffffffu;  // R >= thr, compared from the least significant plane up
#pragma unroll
    for (int k = 0; k < 12; ++k)  // ge = t ? (w & ge) : (w | ge) -- one LOP3 per plane
      asm("lop3.b32 %0, %1, %0, %2, 0xD4;" : "+r"(ge) : "r"(w[k]), "r"(d.tmask[k]));
    return ge;
  }
  // v[j] = keep_j ? v[j] * scale : 0 for 8 consecutive elements whose keep bits are bits 0..7 of `kb`
  __device__ __forceinline__ static void apply8(const DropCfg& d, uint32_t kb, float (&v)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (kb & (1u << j)) ? v[j] * d.scale : 0.f;
  }
  // Warp-cooperative form for kernels in which lane l owns the 8-element chunks c = l + 32 i (i < NCH) of a row of
  // `nchunks` chunks (nchunks % 4 == 0, nchunks <= 128): lanes 0 .. nchunks / 4 - 1 generate one word each, every lane
  // then fetches the byte of each of its chunks.  Must be called by all 32 lanes.  `first_word` = word index of the
  // row's first element.
  template <int NCH>
  __device__ __forceinline__ void keep_bytes_row(const DropCfg& d, uint64_t first_word, int nchunks, int lane,
                                                 uint32_t (&kb)[NCH]) const {
    uint32_t w = 0xffffffffu;
    if (lane < (nchunks >> 2)) w = keep32(d, first_word + lane);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      kb[i] = (__shfl_sync(0xffffffffu, w, (c >> 2) & 31) >> (8 * (c & 3))) & 0xffu;
    }
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
