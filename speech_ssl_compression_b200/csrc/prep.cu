// Data-movement kernels around the GEMMs: weight preparation (fp32 master [+ prune mask] ->
// bf16 operand and its transpose), casts, input masking, masked-frame selection / gather /
// scatter.  All HBM-bound: coalesced 16-byte accesses, grids sized in multiples of the SM count.
#include "mh_b200.h"
#include "mh_common.cuh"

namespace mh {
extern long long g_launches;

// 64 x 64 tile per block (256 threads).  Reads fp32 rows coalesced, writes bf16 rows of dst and,
// through a padded smem tile, bf16 rows of dst_t (the transpose).
__global__ void __launch_bounds__(256)
weight_prep_kernel(const float* __restrict__ src, const uint8_t* __restrict__ mask, __nv_bfloat16* __restrict__ dst,
                   long long ld_dst, __nv_bfloat16* __restrict__ dst_t, long long ld_dst_t, int rows, int cols) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 columns each
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 16 * i, c = c0 + tx * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows && c < cols) {  // cols % 4 == 0
      v = __ldg(reinterpret_cast<const float4*>(src + static_cast<long long>(r) * cols + c));
      if (mask != nullptr) {
        const uint32_t m = *reinterpret_cast<const uint32_t*>(mask + static_cast<long long>(r) * cols + c);
        if ((m & 0xFFu) == 0) v.x = 0.f;
        if ((m & 0xFF00u) == 0) v.y = 0.f;
        if ((m & 0xFF0000u) == 0) v.z = 0.f;
        if ((m & 0xFF000000u) == 0) v.w = 0.f;
      }
      uint2 o = make_uint2(f32x2_to_bf16(v.x, v.y), f32x2_to_bf16(v.z, v.w));
      *reinterpret_cast<uint2*>(dst + static_cast<long long>(r) * ld_dst + c) = o;
    }
    tile[ty + 16 * i][tx * 4 + 0] = __float2bfloat16(v.x);
    tile[ty + 16 * i][tx * 4 + 1] = __float2bfloat16(v.y);
    tile[ty + 16 * i][tx * 4 + 2] = __float2bfloat16(v.z);
    tile[ty + 16 * i][tx * 4 + 3] = __float2bfloat16(v.w);
  }
  if (dst_t == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 16 * i;  // row of dst_t
    const int r = r0 + tx * 4;       // 4 consecutive columns of dst_t
    if (c < cols && r < rows) {      // rows % 4 == 0 checked on the host
      __nv_bfloat16 a = tile[tx * 4 + 0][ty + 16 * i], b = tile[tx * 4 + 1][ty + 16 * i];
      __nv_bfloat16 cc = tile[tx * 4 + 2][ty + 16 * i], d = tile[tx * 4 + 3][ty + 16 * i];
      __nv_bfloat162 lo = __halves2bfloat162(a, b), hi = __halves2bfloat162(cc, d);
      uint2 o = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      *reinterpret_cast<uint2*>(dst_t + static_cast<long long>(c) * ld_dst_t + r) = o;
    }
  }
}

__global__ void bias_prep_kernel(const float* __restrict__ src, const uint8_t* __restrict__ mask,
                                 float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (mask == nullptr || mask[i]) ? src[i] : 0.f;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + i + 4));
    stg128(dst + i, make_uint4(f32x2_to_bf16(a.x, a.y), f32x2_to_bf16(a.z, a.w), f32x2_to_bf16(b.x, b.y),
                               f32x2_to_bf16(b.z, b.w)));
  }
  if (i < n && i + 8 > n)
    for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
}

__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n) {
  long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    float v[8];
    bf16x8_to_f32(ldg128(src + i), v);
    *reinterpret_cast<float4*>(dst + i) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dst + i + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (i < n && i + 8 > n)
    for (long long j = i; j < n; ++j) dst[j] = __bfloat162float(src[j]);
}

// dst(bf16)[r, :] = zero[r] ? 0 : src(f32)[r, :]      cols % 8 == 0
__global__ void mask_rows_kernel(const float* __restrict__ src, const uint8_t* __restrict__ zero_row,
                                 __nv_bfloat16* __restrict__ dst, int rows, int cols) {
  const int cpr = cols >> 3;
  const long long total = static_cast<long long>(rows) * cpr;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cpr), c = static_cast<int>(i % cpr) * 8;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (zero_row == nullptr || !zero_row[r]) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src + static_cast<long long>(r) * cols + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src + static_cast<long long>(r) * cols + c + 4));
      o = make_uint4(f32x2_to_bf16(a.x, a.y), f32x2_to_bf16(a.z, a.w), f32x2_to_bf16(b.x, b.y), f32x2_to_bf16(b.z, b.w));
    }
    stg128(dst + static_cast<long long>(r) * cols + c, o);
  }
}

__global__ void zero_rows_kernel(__nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ zero_row, int rows,
                                 int cols) {
  const int cpr = cols >> 3;
  const long long total = static_cast<long long>(rows) * cpr;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cpr), c = static_cast<int>(i % cpr) * 8;
    if (zero_row[r]) stg128(x + static_cast<long long>(r) * cols + c, make_uint4(0, 0, 0, 0));
  }
}

// Ordered compaction of the selected row indices by a single 1024-thread block (rows <= a few
// hundred thousand): each thread owns a contiguous slice, block-wide exclusive scan of counts.
__global__ void __launch_bounds__(1024)
select_rows_kernel(const uint8_t* __restrict__ sel, int* __restrict__ idx, int* __restrict__ count, int rows) {
  __shared__ int warp_tot[32];
  __shared__ int total_s;
  const int per = (rows + 1023) / 1024;
  const int b = threadIdx.x * per, e = min(rows, b + per);
  int c = 0;
  for (int i = b; i < e; ++i) c += sel[i] != 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_tot[lane] = winc - w;
    if (lane == 31) total_s = winc;
  }
  __syncthreads();
  int pos = warp_tot[warp] + inc - c;
  for (int i = b; i < e; ++i)
    if (sel[i] != 0) idx[pos++] = i;
  const int total = total_s;
  if (threadIdx.x == 0) *count = total;
  for (int i = total + threadIdx.x; i < rows; i += 1024) idx[i] = -1;
}

__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, const int* __restrict__ idx,
                                   __nv_bfloat16* __restrict__ dst, int n_idx, int cols) {
  const int cpr = cols >> 3;
  const long long total = static_cast<long long>(n_idx) * cpr;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cpr), c = static_cast<int>(i % cpr) * 8;
    const int s = idx[r];
    uint4 v = make_uint4(0, 0, 0, 0);
    if (s >= 0) v = ldg128(src + static_cast<long long>(s) * cols + c);
    stg128(dst + static_cast<long long>(r) * cols + c, v);
  }
}

__global__ void scatter_rows_add_kernel(const __nv_bfloat16* __restrict__ src, const int* __restrict__ idx,
                                        __nv_bfloat16* __restrict__ dst, int n_idx, int cols) {
  const int cpr = cols >> 3;
  const long long total = static_cast<long long>(n_idx) * cpr;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cpr), c = static_cast<int>(i % cpr) * 8;
    const int s = idx[r];
    if (s < 0) continue;
    float a[8], b[8];
    bf16x8_to_f32(ldg128(src + static_cast<long long>(r) * cols + c), a);
    __nv_bfloat16* d = dst + static_cast<long long>(s) * cols + c;
    bf16x8_to_f32(*reinterpret_cast<const uint4*>(d), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    stg128(d, f32_to_bf16x8(a));
  }
}

__global__ void gather_labels_kernel(const long long* __restrict__ label, const int* __restrict__ idx,
                                     long long* __restrict__ dst, int n_idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_idx) dst[i] = idx[i] >= 0 ? label[idx[i]] : -100;
}

// y = x * keep(seed, site) * scale, element index = row * cols + col (the GEMM-epilogue indexing): one dropout stream
// word (32 consecutive elements) per thread
__global__ void dropout_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long words,
                                     const DropCfg drop) {
  const DropState dstate(drop);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < words;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint32_t kw = drop.thresh != 0 ? dstate.keep32(drop, static_cast<uint64_t>(i)) : 0xffffffffu;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[8];
      bf16x8_to_f32(ldg128(x + i * 32 + q * 8), v);
      if (drop.thresh != 0) DropState::apply8(drop, kw >> (8 * q), v);
      stg128(y + i * 32 + q * 8, f32_to_bf16x8(v));
    }
  }
}

static int ew_grid(long long work_items, int threads) {
  long long g = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}
}  // namespace mh

using namespace mh;
#define ST reinterpret_cast<cudaStream_t>(stream)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" int mh_weight_prep(const float* src, const uint8_t* mask, void* dst, long long ld_dst, void* dst_t,
                              long long ld_dst_t, int rows, int cols, void* stream) {
  MH_CHECK(rows > 0 && cols > 0 && cols % 4 == 0, "weight_prep: cols must be a multiple of 4 (%d x %d)", rows, cols);
  MH_CHECK(dst_t == nullptr || rows % 4 == 0, "weight_prep: rows must be a multiple of 4 for the transposed copy");
  weight_prep_kernel<<<dim3((cols + 63) / 64, (rows + 63) / 64), 256, 0, ST>>>(src, mask, BF(dst), ld_dst, BF(dst_t),
                                                                                ld_dst_t, rows, cols);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

extern "C" int mh_bias_prep(const float* src, const uint8_t* mask, float* dst, int n, void* stream) {
  bias_prep_kernel<<<(n + 255) / 256, 256, 0, ST>>>(src, mask, dst, n);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

extern "C" int mh_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
  MH_CHECK((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "cast: alignment");
  cast_f32_bf16_kernel<<<ew_grid((n + 7) / 8, 256), 256, 0, ST>>>(src, BF(dst), n);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

extern "C" int mh_cast_bf16_to_f32(const void* src, float* dst, long long n, void* stream) {
  MH_CHECK((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "cast: alignment");
  cast_bf16_f32_kernel<<<ew_grid((n + 7) / 8, 256), 256, 0, ST>>>(CBF(src), dst, n);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

extern "C" int mh_mask_rows_f32_to_bf16(const float* src, const uint8_t* zero_row, void* dst, int rows, int cols,
                                        void* stream) {
  MH_CHECK(cols % 8 == 0, "mask_rows: cols %% 8");
  mask_rows_kernel<<<ew_grid(static_cast<long long>(rows) * (cols / 8), 256), 256, 0, ST>>>(src, zero_row, BF(dst), rows,
                                                                                           cols);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

extern "C" int mh_zero_rows_bf16(void* x, const uint8_t* zero_row, int rows, int cols, void* stream) {
  MH_CHECK(cols % 8 == 0, "zero_rows: cols %% 8");
  zero_rows_kernel<<<ew_grid(static_cast<long long>(rows) * (cols / 8), 256), 256, 0, ST>>>(BF(x), zero_row, rows, cols);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

extern "C" int mh_select_rows(const uint8_t* sel, int* idx, int* count, int rows, void* stream) {
  select_rows_kernel<<<1, 1024, 0, ST>>>(sel, idx, count, rows);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

extern "C" int mh_gather_rows(const void* src, const int* idx, void* dst, int n_idx, int cols, void* stream) {
  MH_CHECK(cols % 8 == 0, "gather_rows: cols %% 8");
  if (n_idx == 0) return 0;
  gather_rows_kernel<<<ew_grid(static_cast<long long>(n_idx) * (cols / 8), 256), 256, 0, ST>>>(CBF(src), idx, BF(dst),
                                                                                              n_idx, cols);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

extern "C" int mh_scatter_rows_add(const void* src, const int* idx, void* dst, int n_idx, int cols, void* stream) {
  MH_CHECK(cols % 8 == 0, "scatter_rows: cols %% 8");
  if (n_idx == 0) return 0;
  scatter_rows_add_kernel<<<ew_grid(static_cast<long long>(n_idx) * (cols / 8), 256), 256, 0, ST>>>(CBF(src), idx,
                                                                                                   BF(dst), n_idx, cols);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

extern "C" int mh_gather_labels(const long long* label, const int* idx, long long* dst, int n_idx, void* stream) {
  if (n_idx == 0) return 0;
  gather_labels_kernel<<<(n_idx + 255) / 256, 256, 0, ST>>>(label, idx, dst, n_idx);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

extern "C" int mh_dropout_apply(const void* x, void* y, int rows, int cols, float p_drop, uint64_t seed, uint32_t site,
                                void* stream) {
  MH_CHECK(cols % 32 == 0 && p_drop > 0.f, "dropout_apply: cols %% 32 (one dropout stream word = 32 elements) and p > 0 required");
  const long long words = static_cast<long long>(rows) * (cols / 32);
  dropout_apply_kernel<<<ew_grid(words, 256), 256, 0, ST>>>(CBF(x), BF(y), words, make_drop(p_drop, seed, site));
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}
