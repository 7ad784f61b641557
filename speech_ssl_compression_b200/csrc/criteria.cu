// Fused criterion kernels: masked-frame cross-entropy, KD (CE + KL) and the per-frame
// L1 + cosine distillation loss.  One warp per row, 16-byte loads, fp32 math, warp-shuffle
// reductions, one atomic per warp into the accumulators.  HBM-bound (each logit read once
// per pass).
#include "mh_b200.h"
#include "mh_common.cuh"

namespace mh {
extern long long g_launches;
constexpr int CR_WARPS = 8;

// loads a row of `cols` bf16 into v[NCH][8] (chunk c = lane + 32 i), -inf / 0 padding
template <int NCH>
__device__ __forceinline__ void load_row(const __nv_bfloat16* p, int nchunks, int lane, float (&v)[NCH][8], float pad) {
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = lane + 32 * i;
    if (c < nchunks) bf16x8_to_f32(ldg128(p + c * 8), v[i]);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = pad;
    }
  }
}
template <int NCH>
__device__ __forceinline__ float row_max(const float (&v)[NCH][8]) {
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < NCH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) m = fmaxf(m, v[i][j]);
  return warp_max(m);
}
// sum exp((v - m) * inv_t)
template <int NCH>
__device__ __forceinline__ float row_sumexp(const float (&v)[NCH][8], float m, float inv_t) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += __expf((v[i][j] - m) * inv_t);
  return warp_sum(s);
}
template <int NCH>
__device__ __forceinline__ float pick(const float (&v)[NCH][8], int lane, int label) {
  float x = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if ((lane + 32 * i) * 8 + j == label) x = v[i][j];
  return warp_sum(x);
}

template <int NCH>
__global__ void __launch_bounds__(CR_WARPS * 32)
ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, const long long* __restrict__ labels,
              const int* __restrict__ n_valid, float* __restrict__ row_loss, float* __restrict__ acc, int n_rows,
              int n_class) {
  const int lane = threadIdx.x & 31;
  const int limit = n_valid ? min(n_rows, *n_valid) : n_rows;
  const int nchunks = n_class >> 3;
  float loss_sum = 0.f, cnt = 0.f;
  for (int row = blockIdx.x * CR_WARPS + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * CR_WARPS) {
    float l = 0.f;
    const long long lab = row < limit ? labels[row] : -100;
    if (lab >= 0) {
      float v[NCH][8];
      load_row<NCH>(logits + static_cast<long long>(row) * n_class, nchunks, lane, v, -INFINITY);
      const float m = row_max<NCH>(v);
      const float lse = m + __logf(row_sumexp<NCH>(v, m, 1.f));
      l = lse - pick<NCH>(v, lane, static_cast<int>(lab));
      loss_sum += l;
      cnt += 1.f;
    }
    if (row_loss && lane == 0) row_loss[row] = l;
  }
  if (lane == 0 && cnt > 0.f) {
    atomicAdd(acc + 0, loss_sum);
    atomicAdd(acc + 1, cnt);
  }
}

template <int NCH>
__global__ void __launch_bounds__(CR_WARPS * 32)
ce_bwd_kernel(const __nv_bfloat16* __restrict__ logits, const long long* __restrict__ labels,
              const int* __restrict__ n_valid, const float* __restrict__ grad_scale,
              __nv_bfloat16* __restrict__ dlogits, int n_rows, int n_class) {
  const int lane = threadIdx.x & 31;
  const int limit = n_valid ? min(n_rows, *n_valid) : n_rows;
  const int nchunks = n_class >> 3;
  const float gs = *grad_scale;
  for (int row = blockIdx.x * CR_WARPS + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * CR_WARPS) {
    const long long lab = row < limit ? labels[row] : -100;
    __nv_bfloat16* out = dlogits + static_cast<long long>(row) * n_class;
    if (lab < 0) {
#pragma unroll
      for (int i = 0; i < NCH; ++i)
        if (lane + 32 * i < nchunks) stg128(out + (lane + 32 * i) * 8, make_uint4(0, 0, 0, 0));
      continue;
    }
    float v[NCH][8];
    load_row<NCH>(logits + static_cast<long long>(row) * n_class, nchunks, lane, v, -INFINITY);
    const float m = row_max<NCH>(v);
    const float inv = gs / row_sumexp<NCH>(v, m, 1.f);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = __expf(v[i][j] - m) * inv;
          if (c * 8 + j == lab) o[j] -= gs;
        }
        stg128(out + c * 8, f32_to_bf16x8(o));
      }
    }
  }
}

// KD: per row  hard = CE(s, label), teacher CE, soft = sum_c pt (log pt - log ps) with
// temperature T on both.
template <int NCH>
__global__ void __launch_bounds__(CR_WARPS * 32)
kd_fwd_kernel(const __nv_bfloat16* __restrict__ s_logits, const __nv_bfloat16* __restrict__ t_logits,
              const long long* __restrict__ labels, const int* __restrict__ n_valid, float inv_t,
              float* __restrict__ acc, int n_rows, int n_class) {
  const int lane = threadIdx.x & 31;
  const int limit = n_valid ? min(n_rows, *n_valid) : n_rows;
  const int nchunks = n_class >> 3;
  float a_hard = 0.f, a_cnt = 0.f, a_soft = 0.f, a_tce = 0.f, a_rows = 0.f;
  for (int row = blockIdx.x * CR_WARPS + (threadIdx.x >> 5); row < limit; row += gridDim.x * CR_WARPS) {
    const long long lab = labels[row];
    if (lab < 0) continue;  // padded slot of a statically shaped row list
    float s[NCH][8], t[NCH][8];
    load_row<NCH>(s_logits + static_cast<long long>(row) * n_class, nchunks, lane, s, -INFINITY);
    load_row<NCH>(t_logits + static_cast<long long>(row) * n_class, nchunks, lane, t, -INFINITY);
    const float ms = row_max<NCH>(s), mt = row_max<NCH>(t);
    {
      a_hard += ms + __logf(row_sumexp<NCH>(s, ms, 1.f)) - pick<NCH>(s, lane, static_cast<int>(lab));
      a_tce += mt + __logf(row_sumexp<NCH>(t, mt, 1.f)) - pick<NCH>(t, lane, static_cast<int>(lab));
      a_cnt += 1.f;
    }
    const float zs = row_sumexp<NCH>(s, ms, inv_t), zt = row_sumexp<NCH>(t, mt, inv_t);
    const float lzs = __logf(zs), lzt = __logf(zt);
    float kl = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
      if (lane + 32 * i < nchunks) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float ls = (s[i][j] - ms) * inv_t - lzs, lt = (t[i][j] - mt) * inv_t - lzt;
          kl += __expf(lt) * (lt - ls);
        }
      }
    a_soft += warp_sum(kl);
    a_rows += 1.f;
  }
  if (lane == 0 && a_rows > 0.f) {
    atomicAdd(acc + 0, a_hard);
    atomicAdd(acc + 1, a_cnt);
    atomicAdd(acc + 2, a_soft);
    atomicAdd(acc + 3, a_tce);
    atomicAdd(acc + 4, a_rows);
  }
}

template <int NCH>
__global__ void __launch_bounds__(CR_WARPS * 32)
kd_bwd_kernel(const __nv_bfloat16* __restrict__ s_logits, const __nv_bfloat16* __restrict__ t_logits,
              const long long* __restrict__ labels, const int* __restrict__ n_valid, float inv_t,
              const float* __restrict__ w_hard, const float* __restrict__ w_soft, __nv_bfloat16* __restrict__ dlogits,
              int n_rows, int n_class) {
  const int lane = threadIdx.x & 31;
  const int limit = n_valid ? min(n_rows, *n_valid) : n_rows;
  const int nchunks = n_class >> 3;
  const float wh = *w_hard, ws = *w_soft * inv_t;
  for (int row = blockIdx.x * CR_WARPS + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * CR_WARPS) {
    __nv_bfloat16* out = dlogits + static_cast<long long>(row) * n_class;
    if (row >= limit || labels[row] < 0) {
#pragma unroll
      for (int i = 0; i < NCH; ++i)
        if (lane + 32 * i < nchunks) stg128(out + (lane + 32 * i) * 8, make_uint4(0, 0, 0, 0));
      continue;
    }
    float s[NCH][8], t[NCH][8];
    load_row<NCH>(s_logits + static_cast<long long>(row) * n_class, nchunks, lane, s, -INFINITY);
    load_row<NCH>(t_logits + static_cast<long long>(row) * n_class, nchunks, lane, t, -INFINITY);
    const float ms = row_max<NCH>(s), mt = row_max<NCH>(t);
    const long long lab = labels[row];
    const float h = lab >= 0 ? wh : 0.f;
    const float i1 = h / row_sumexp<NCH>(s, ms, 1.f);
    const float isT = ws / row_sumexp<NCH>(s, ms, inv_t), itT = ws / row_sumexp<NCH>(t, mt, inv_t);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = __expf(s[i][j] - ms) * i1 + __expf((s[i][j] - ms) * inv_t) * isT - __expf((t[i][j] - mt) * inv_t) * itT;
          if (c * 8 + j == lab) o[j] -= h;
        }
        stg128(out + c * 8, f32_to_bf16x8(o));
      }
    }
  }
}

template <int NCH>
__global__ void __launch_bounds__(CR_WARPS * 32)
l1cos_fwd_kernel(const __nv_bfloat16* __restrict__ pred, const __nv_bfloat16* __restrict__ tgt,
                 float* __restrict__ acc, int rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int nchunks = cols >> 3;
  float a_l1 = 0.f, a_cos = 0.f;
  for (int row = blockIdx.x * CR_WARPS + (threadIdx.x >> 5); row < rows; row += gridDim.x * CR_WARPS) {
    float p[NCH][8], t[NCH][8];
    load_row<NCH>(pred + static_cast<long long>(row) * cols, nchunks, lane, p, 0.f);
    load_row<NCH>(tgt + static_cast<long long>(row) * cols, nchunks, lane, t, 0.f);
    float l1 = 0.f, pt = 0.f, pp = 0.f, tt = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        l1 += fabsf(p[i][j] - t[i][j]);
        pt += p[i][j] * t[i][j];
        pp += p[i][j] * p[i][j];
        tt += t[i][j] * t[i][j];
      }
    l1 = warp_sum(l1); pt = warp_sum(pt); pp = warp_sum(pp); tt = warp_sum(tt);
    const float cosv = pt / (fmaxf(sqrtf(pp), 1e-8f) * fmaxf(sqrtf(tt), 1e-8f));
    a_l1 += l1;
    a_cos += log1pf(__expf(-cosv));  // -logsigmoid(cos)
  }
  if (lane == 0) {
    atomicAdd(acc + 0, a_l1);
    atomicAdd(acc + 1, a_cos);
  }
}

template <int NCH>
__global__ void __launch_bounds__(CR_WARPS * 32)
l1cos_bwd_kernel(const __nv_bfloat16* __restrict__ pred, const __nv_bfloat16* __restrict__ tgt,
                 const float* __restrict__ w_l1, const float* __restrict__ w_cos, __nv_bfloat16* __restrict__ dpred,
                 int rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int nchunks = cols >> 3;
  const float wl = *w_l1, wc = *w_cos;
  for (int row = blockIdx.x * CR_WARPS + (threadIdx.x >> 5); row < rows; row += gridDim.x * CR_WARPS) {
    float p[NCH][8], t[NCH][8];
    load_row<NCH>(pred + static_cast<long long>(row) * cols, nchunks, lane, p, 0.f);
    load_row<NCH>(tgt + static_cast<long long>(row) * cols, nchunks, lane, t, 0.f);
    float pt = 0.f, pp = 0.f, tt = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        pt += p[i][j] * t[i][j];
        pp += p[i][j] * p[i][j];
        tt += t[i][j] * t[i][j];
      }
    pt = warp_sum(pt); pp = warp_sum(pp); tt = warp_sum(tt);
    const float np = fmaxf(sqrtf(pp), 1e-8f), nt = fmaxf(sqrtf(tt), 1e-8f);
    const float cosv = pt / (np * nt);
    // d/dcos [-logsigmoid(cos)] = -sigmoid(-cos);  dcos/dp = t/(np nt) - cos * p / np^2
    const float g = -wc / (1.f + __expf(cosv));
    const float ca = g / (np * nt), cb = g * cosv / (np * np);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunks) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = p[i][j] - t[i][j];
          o[j] = wl * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) + ca * t[i][j] - cb * p[i][j];
        }
        stg128(dpred + static_cast<long long>(row) * cols + c * 8, f32_to_bf16x8(o));
      }
    }
  }
}

// loss = w * acc[0] / acc[1]; grad_scale = w / acc[1]  (single thread; keeps the normaliser on
// the device so CUDA graphs / all-reduced counts need no host round trip)
__global__ void ce_finalize_kernel(const float* acc, float weight, float* loss, float* grad_scale) {
  const float n = fmaxf(acc[1], 1.f);
  if (loss) *loss = weight * acc[0] / n;
  if (grad_scale) *grad_scale = weight / n;
}

__global__ void kd_finalize_kernel(const float* acc, float alpha, float* out, float* w_hard, float* w_soft) {
  const float cnt = fmaxf(acc[1], 1.f), rows = fmaxf(acc[4], 1.f);
  const float hard = acc[0] / cnt, soft = acc[2] / rows;
  if (out) {
    out[0] = hard * (1.f - alpha) + soft * alpha;
    out[1] = hard;
    out[2] = soft;
    out[3] = acc[3] / cnt;
  }
  if (w_hard) *w_hard = (1.f - alpha) / cnt;
  if (w_soft) *w_soft = alpha / rows;
}

template <typename F>
static int dispatch_cols(int cols, const char* what, F&& f) {
  if (cols % 8 != 0 || cols <= 0) { set_error("%s: columns must be a positive multiple of 8 (got %d)", what, cols); return 1; }
  const int nch = (cols / 8 + 31) / 32;
  switch (nch) {
    case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 3: return f(std::integral_constant<int, 3>());
    case 4: return f(std::integral_constant<int, 4>());
    default: set_error("%s: at most 1024 columns supported (got %d)", what, cols); return 1;
  }
}
static int rows_grid(int rows) {
  int g = (rows + CR_WARPS - 1) / CR_WARPS;
  const int cap = sm_count() * 8;
  return g > cap ? cap : (g < 1 ? 1 : g);
}
}  // namespace mh

using namespace mh;
#define ST reinterpret_cast<cudaStream_t>(stream)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define LAUNCHED()      \
  MH_LAUNCH_CHECK();    \
  ++g_launches;         \
  return 0

extern "C" int mh_ce_fwd(const void* logits, const long long* labels, const int* n_valid, float* row_loss, float* acc,
                         int n_rows, int n_class, void* stream) {
  if (n_rows == 0) return 0;
  return dispatch_cols(n_class, "ce_fwd", [&](auto nch) {
    ce_fwd_kernel<decltype(nch)::value><<<rows_grid(n_rows), CR_WARPS * 32, 0, ST>>>(CBF(logits), labels, n_valid, row_loss,
                                                                                    acc, n_rows, n_class);
    LAUNCHED();
  });
}
extern "C" int mh_ce_bwd(const void* logits, const long long* labels, const int* n_valid, const float* grad_scale,
                         void* dlogits, int n_rows, int n_class, void* stream) {
  if (n_rows == 0) return 0;
  return dispatch_cols(n_class, "ce_bwd", [&](auto nch) {
    ce_bwd_kernel<decltype(nch)::value><<<rows_grid(n_rows), CR_WARPS * 32, 0, ST>>>(CBF(logits), labels, n_valid, grad_scale,
                                                                                    BF(dlogits), n_rows, n_class);
    LAUNCHED();
  });
}
extern "C" int mh_ce_finalize(const float* acc, float weight, float* loss, float* grad_scale, void* stream) {
  ce_finalize_kernel<<<1, 1, 0, ST>>>(acc, weight, loss, grad_scale);
  LAUNCHED();
}
extern "C" int mh_kd_finalize(const float* acc, float alpha, float* out, float* w_hard, float* w_soft, void* stream) {
  kd_finalize_kernel<<<1, 1, 0, ST>>>(acc, alpha, out, w_hard, w_soft);
  LAUNCHED();
}
extern "C" int mh_kd_fwd(const void* s_logits, const void* t_logits, const long long* labels, const int* n_valid, float T,
                         float* acc, int n_rows, int n_class, void* stream) {
  if (n_rows == 0) return 0;
  return dispatch_cols(n_class, "kd_fwd", [&](auto nch) {
    kd_fwd_kernel<decltype(nch)::value><<<rows_grid(n_rows), CR_WARPS * 32, 0, ST>>>(CBF(s_logits), CBF(t_logits), labels,
                                                                                    n_valid, 1.f / T, acc, n_rows, n_class);
    LAUNCHED();
  });
}
extern "C" int mh_kd_bwd(const void* s_logits, const void* t_logits, const long long* labels, const int* n_valid, float T,
                         const float* w_hard, const float* w_soft, void* dlogits, int n_rows, int n_class, void* stream) {
  if (n_rows == 0) return 0;
  return dispatch_cols(n_class, "kd_bwd", [&](auto nch) {
    kd_bwd_kernel<decltype(nch)::value><<<rows_grid(n_rows), CR_WARPS * 32, 0, ST>>>(
        CBF(s_logits), CBF(t_logits), labels, n_valid, 1.f / T, w_hard, w_soft, BF(dlogits), n_rows, n_class);
    LAUNCHED();
  });
}
extern "C" int mh_l1cos_fwd(const void* pred, const void* target, float* acc, int rows, int cols, void* stream) {
  if (rows == 0) return 0;
  return dispatch_cols(cols, "l1cos_fwd", [&](auto nch) {
    l1cos_fwd_kernel<decltype(nch)::value><<<rows_grid(rows), CR_WARPS * 32, 0, ST>>>(CBF(pred), CBF(target), acc, rows, cols);
    LAUNCHED();
  });
}
extern "C" int mh_l1cos_bwd(const void* pred, const void* target, const float* w_l1, const float* w_cos, void* dpred,
                            int rows, int cols, void* stream) {
  if (rows == 0) return 0;
  return dispatch_cols(cols, "l1cos_bwd", [&](auto nch) {
    l1cos_bwd_kernel<decltype(nch)::value><<<rows_grid(rows), CR_WARPS * 32, 0, ST>>>(CBF(pred), CBF(target), w_l1, w_cos,
                                                                                     BF(dpred), rows, cols);
    LAUNCHED();
  });
}
