// LayerNorm forward / backward and bias-gradient column sums.  HBM-bound kernels: one warp
// per row, 16-byte loads, fp32 statistics, warp-shuffle reductions, no shared-memory staging
// of the row (each element is touched once).
#include "mh_b200.h"
#define MH_PDL_FAMILY 4
#include "mh_common.cuh"

namespace mh {
extern long long g_launches;

constexpr int LN_WARPS = 8;

// NCH = number of 8-element chunks per lane (cols <= NCH * 256)
template <int NCH>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows,
              int cols, float eps, DropCfg drop) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * LN_WARPS;
  const int nchunks = cols >> 3;
  for (int row = warp_global; row < rows; row += nwarps) {
    const __nv_bfloat16* xr = x + static_cast<long long>(row) * cols;
    float v[NCH][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        bf16x8_to_f32(ldg128(xr + c * 8), v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      }
    }
    const float mean = warp_sum(s) / cols;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mean;
          sq += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / cols + eps);
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
    __nv_bfloat16* yr = y + static_cast<long long>(row) * cols;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c * 8 + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
        if (drop.thresh != 0) {
          drop_apply8(drop, static_cast<uint64_t>(row) * nchunks + c, o);
        }
        stg128(yr + c * 8, f32_to_bf16x8(o));
      }
    }
  }
}

// Backward.  dy_eff = dy (* keep mask of the output dropout `din`, if the forward dropped the
// LN output).  dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)).
// Also emits dx_drop = dx * keep mask of `dout` (the dropout that sat on the GEMM output
// feeding this LayerNorm's input) so the dgrad/wgrad GEMMs can consume it directly.
#ifndef LN_BWD_MINB
#define LN_BWD_MINB 2
#endif
#ifndef LN_BWD_GRID_MULT
#define LN_BWD_GRID_MULT 2
#endif
template <int NCH>
__global__ void __launch_bounds__(LN_WARPS * 32, LN_BWD_MINB)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
              const float* __restrict__ gamma, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dx_drop, float* __restrict__ dgamma,
              float* __restrict__ dbeta, int rows, int cols, DropCfg din, DropCfg dout) {
  pdl_prologue();
  __shared__ float red[LN_WARPS][32 * 8 + 1];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warp_global = blockIdx.x * LN_WARPS + warp;
  const int nwarps = gridDim.x * LN_WARPS;
  const int nchunks = cols >> 3;
  float dg[NCH][8], db[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) dg[i][j] = db[i][j] = 0.f;

  // The row is kept as the raw bf16 it was loaded as (and unpacked twice) instead of as fp32 products: that leaves
  // the registers to have the NEXT row's loads in flight while this one is reduced and written -- the kernel is
  // HBM-bound and a load -> shuffle-reduce -> store sequence per row left the memory pipe idle half of the time.
  uint4 nd[NCH], nx[NCH];
  float nmean = 0.f, nrstd = 0.f;
  auto fetch = [&](int row) {
    const long long off = static_cast<long long>(row) * cols;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        nd[i] = ldg128(dy + off + c * 8);
        nx[i] = ldg128(x + off + c * 8);
      }
    }
    nmean = __ldg(mean_in + row);
    nrstd = __ldg(rstd_in + row);
  };
  if (warp_global < rows) fetch(warp_global);
  for (int row = warp_global; row < rows; row += nwarps) {
    const long long off = static_cast<long long>(row) * cols;
    uint4 rd[NCH], rx[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) { rd[i] = nd[i]; rx[i] = nx[i]; }
    const float mean = nmean, rstd = nrstd;
    if (row + nwarps < rows) fetch(row + nwarps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float d[8], xv[8];
        bf16x8_to_f32(rd[i], d);
        bf16x8_to_f32(rx[i], xv);
        if (din.thresh != 0) {
          drop_apply8(din, static_cast<uint64_t>(row) * nchunks + c, d);
        }
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c * 8 + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          const float gy = d[j] * g[j];
          s1 += gy;
          s2 += gy * xh;
          dg[i][j] += d[j] * xh;
          db[i][j] += d[j];
        }
      }
    }
    s1 = warp_sum(s1) / cols;
    s2 = warp_sum(s2) / cols;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float d[8], xv[8], o[8];
        bf16x8_to_f32(rd[i], d);
        bf16x8_to_f32(rx[i], xv);
        if (din.thresh != 0) {
          drop_apply8(din, static_cast<uint64_t>(row) * nchunks + c, d);  // (regenerated: one LN instance per step)
        }
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c * 8 + 4));
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (d[j] * g[j] - s1 - ((xv[j] - mean) * rstd) * s2);
        stg128(dx + off + c * 8, f32_to_bf16x8(o));
        if (dx_drop != nullptr) {
          if (dout.thresh != 0) {
            drop_apply8(dout, static_cast<uint64_t>(row) * nchunks + c, o);
          }
          stg128(dx_drop + off + c * 8, f32_to_bf16x8(o));
        }
      }
    }
  }
  // block reduction of the per-warp partial dgamma / dbeta, then one atomic per column
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = pass == 0 ? dg[i][j] : db[i][j];
      __syncthreads();
      const int c = threadIdx.x;  // 256 threads <-> 256 columns of this chunk group
      const int col = (32 * i) * 8 + c;
      if (col < cols) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) t += red[w][c];
        atomicAdd((pass == 0 ? dgamma : dbeta) + col, t);
      }
    }
  }
}

// out[n] += sum_m x[m, n].  Block = 8 warps over a slab of 256 columns and a slice of rows.
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ld, float* __restrict__ out, int rows, int cols,
              int rows_per_block) {
  pdl_prologue();
  __shared__ float red[8][32 * 8 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col0 = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col0 < cols) {
    // four independent 16-byte loads in flight per lane: with one the short row slices were pure load latency
    const __nv_bfloat16* xp = x + col0;
    int r = r0 + warp;
    for (; r + 24 < r1; r += 32) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = ldg128(xp + static_cast<long long>(r + 8 * u) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[8];
        bf16x8_to_f32(q[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    }
    for (; r < r1; r += 8) {
      float v[8];
      bf16x8_to_f32(ldg128(xp + static_cast<long long>(r) * ld), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col < cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(out + col, t);
  }
}

template <typename F>
static int dispatch_nch(int cols, F&& f) {
  const int nch = (cols / 8 + 31) / 32;
  switch (nch) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 3: return f(std::integral_constant<int, 3>());
    case 4: return f(std::integral_constant<int, 4>());
    default: set_error("LayerNorm supports up to 1024 columns (got %d)", cols); return 1;
  }
}
}  // namespace mh

using namespace mh;

extern "C" int mh_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                float* rstd, int rows, int cols, float eps, float p_drop, uint64_t seed,
                                uint32_t site, void* stream) {
  MH_CHECK(rows > 0 && cols > 0 && cols % 8 == 0, "layernorm: bad shape %d x %d", rows, cols);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = min((rows + LN_WARPS - 1) / LN_WARPS, sm_count() * 8);
  const DropCfg d = make_drop(p_drop, seed, site);
  return dispatch_nch(cols, [&](auto nch) {
    MH_CUDA(launch_pdl(ln_fwd_kernel<decltype(nch)::value>, dim3(grid), dim3(LN_WARPS * 32), 0, st,
                       reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta, reinterpret_cast<__nv_bfloat16*>(y), mean, rstd,
                       rows, cols, eps, d));
    ++g_launches;
    return 0;
  });
}

extern "C" int mh_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                                const float* rstd, void* dx, void* dx_drop, float* dgamma, float* dbeta, int rows,
                                int cols, float p_in, uint64_t seed_in, uint32_t site_in, float p_out,
                                uint64_t seed_out, uint32_t site_out, void* stream) {
  MH_CHECK(rows > 0 && cols > 0 && cols % 8 == 0, "layernorm_bwd: bad shape %d x %d", rows, cols);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = min((rows + LN_WARPS - 1) / LN_WARPS, sm_count() * LN_BWD_GRID_MULT);
  const DropCfg din = make_drop(p_in, seed_in, site_in), dout = make_drop(p_out, seed_out, site_out);
  return dispatch_nch(cols, [&](auto nch) {
    MH_CUDA(launch_pdl(ln_bwd_kernel<decltype(nch)::value>, dim3(grid), dim3(LN_WARPS * 32), 0, st,
                       reinterpret_cast<const __nv_bfloat16*>(dy), reinterpret_cast<const __nv_bfloat16*>(x), gamma, mean, rstd,
                       reinterpret_cast<__nv_bfloat16*>(dx), reinterpret_cast<__nv_bfloat16*>(dx_drop), dgamma, dbeta, rows,
                       cols, din, dout));
    ++g_launches;
    return 0;
  });
}

extern "C" int mh_colsum(const void* x, long long ld, float* out, int rows, int cols, void* stream) {
  MH_CHECK(rows > 0 && cols > 0 && cols % 8 == 0 && ld % 8 == 0, "colsum: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int gx = (cols + 255) / 256;
  int gy = (sm_count() * 4 + gx - 1) / gx;
  int rpb = (rows + gy - 1) / gy;
  if (rpb < 64) rpb = 64;
  gy = (rows + rpb - 1) / rpb;
  MH_CUDA(launch_pdl(colsum_kernel, dim3(gx, gy), dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(x), ld, out, rows,
                     cols, rpb));
  ++g_launches;
  return 0;
}
