// Shared helpers for all kernels: error plumbing for the C ABI, the counter-based dropout
// RNG (Philox4x32-10), fp32 erf-GELU, small vector load/store helpers.
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

// ---- Philox4x32-10 ----
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0(static_cast<uint32_t>(seed)), k1(static_cast<uint32_t>(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint64_t ctr, uint32_t stream) const {
    uint32_t c0 = static_cast<uint32_t>(ctr), c1 = static_cast<uint32_t>(ctr >> 32), c2 = stream, c3 = 0x9E3779B9u;
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// Dropout over a logical element stream.  One Philox call covers 8 consecutive elements
// (eight 16-bit lanes); element e is kept iff lane(e) >= thresh, thresh = round(p * 65536).
// The forward and the backward regenerate identical decisions from (seed, site, index).
struct DropCfg {
  uint64_t seed;
  const unsigned long long* offset;  // optional device counter added to the seed (CUDA-graph replays)
  uint32_t site;    // unique per dropout site (layer * 8 + site id)
  uint32_t thresh;  // 0 => dropout disabled
  float scale;      // 1 / (1 - p)
};
const unsigned long long* dropout_offset_ptr();
__host__ inline DropCfg make_drop(float p, uint64_t seed, uint32_t site) {
  DropCfg d;
  d.seed = seed;
  d.offset = dropout_offset_ptr();
  d.site = site;
  d.thresh = p > 0.f ? static_cast<uint32_t>(p * 65536.0f + 0.5f) : 0u;
  d.scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  return d;
}
// keep-mask bits for the 8 elements [8*group, 8*group+8)
__device__ __forceinline__ uint32_t drop_keep8(const DropCfg& d, uint64_t group) {
  const uint64_t seed = d.seed + (d.offset != nullptr ? 0x9E3779B97F4A7C15ull * __ldg(d.offset) : 0ull);
  uint4 r = Philox(seed)(group, d.site);
  uint32_t m = 0;
  m |= ((r.x & 0xFFFFu) >= d.thresh) << 0;
  m |= ((r.x >> 16) >= d.thresh) << 1;
  m |= ((r.y & 0xFFFFu) >= d.thresh) << 2;
  m |= ((r.y >> 16) >= d.thresh) << 3;
  m |= ((r.z & 0xFFFFu) >= d.thresh) << 4;
  m |= ((r.z >> 16) >= d.thresh) << 5;
  m |= ((r.w & 0xFFFFu) >= d.thresh) << 6;
  m |= ((r.w >> 16) >= d.thresh) << 7;
  return m;
}

// ---- math ----
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
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
