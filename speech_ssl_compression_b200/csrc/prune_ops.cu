// Pruning-object kernels (cold path, run at prune events only):
//   * exact k-th smallest |w| over a list of fp32 tensors by 3-pass radix select on the bit
//     patterns (11 + 11 + 10 bits), shared-memory histograms;
//   * threshold mask application with the deterministic tie rule "lowest flat index wins";
//   * fp64 L1 row / column sums (head and FFN-row scores).
#include "mh_b200.h"
#include "mh_common.cuh"

namespace mh {
extern long long g_launches;

// workspace layout (u64): [0,2048) histogram | [2048] prefix bits | [2049] prefix mask |
//                         [2050] k remaining | [2051] count below prefix
constexpr int WS_HIST = 0, WS_PREFIX = 2048, WS_PMASK = 2049, WS_KREM = 2050, WS_BELOW = 2051;

__global__ void __launch_bounds__(512)
radix_hist_kernel(const float* __restrict__ w, long long n, int shift, int bits, unsigned long long* __restrict__ ws) {
  __shared__ unsigned int hist[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const uint32_t prefix = static_cast<uint32_t>(ws[WS_PREFIX]), pmask = static_cast<uint32_t>(ws[WS_PMASK]);
  const uint32_t dmask = (1u << bits) - 1u;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint32_t u = __float_as_uint(w[i]) & 0x7FFFFFFFu;
    if ((u & pmask) == prefix) atomicAdd(&hist[(u >> shift) & dmask], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x)
    if (hist[i]) atomicAdd(&ws[WS_HIST + i], static_cast<unsigned long long>(hist[i]));
}

// single thread: locate the digit bucket holding the k-th element, extend the prefix
__global__ void radix_pick_kernel(unsigned long long* ws, int shift, int bits, unsigned long long* result, int last) {
  unsigned long long k = ws[WS_KREM], below = ws[WS_BELOW], run = 0;
  const int nb = 1 << bits;
  int d = 0;
  for (; d < nb; ++d) {
    const unsigned long long h = ws[WS_HIST + d];
    if (run + h >= k) break;
    run += h;
  }
  if (d == nb) d = nb - 1;
  const unsigned long long in_bucket = ws[WS_HIST + d];
  ws[WS_PREFIX] |= static_cast<unsigned long long>(d) << shift;
  ws[WS_PMASK] |= static_cast<unsigned long long>((1u << bits) - 1u) << shift;
  ws[WS_KREM] = k - run;
  ws[WS_BELOW] = below + run;
  for (int i = 0; i < 2048; ++i) ws[WS_HIST + i] = 0;
  if (last) {
    result[0] = ws[WS_PREFIX];   // bit pattern of the k-th smallest |w|
    result[1] = ws[WS_BELOW];    // elements strictly below it
    result[2] = in_bucket;       // elements equal to it
  }
}

__global__ void radix_init_kernel(unsigned long long* ws, unsigned long long k) {
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) ws[WS_HIST + i] = 0;
  if (threadIdx.x == 0) {
    ws[WS_PREFIX] = 0; ws[WS_PMASK] = 0; ws[WS_KREM] = k; ws[WS_BELOW] = 0;
  }
}

// One block walks the tensor in flat order so that ties at the threshold are resolved by
// position.  tie_counter (device, carried across tensors) counts ties already pruned.
__global__ void __launch_bounds__(1024)
apply_threshold_kernel(const float* __restrict__ w, uint8_t* __restrict__ mask, long long n,
                       const unsigned long long* __restrict__ result, unsigned long long k,
                       unsigned long long* __restrict__ tie_counter) {
  __shared__ unsigned long long taken_s;
  __shared__ int warp_cnt[32];
  const uint32_t thr = static_cast<uint32_t>(result[0]);
  const unsigned long long n_ties = k - result[1];  // ties that must be pruned in total
  if (threadIdx.x == 0) taken_s = *tie_counter;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long base = 0; base < n; base += 1024) {
    const long long i = base + threadIdx.x;
    uint32_t u = 0xFFFFFFFFu;
    if (i < n) u = __float_as_uint(w[i]) & 0x7FFFFFFFu;
    const bool tie = (i < n) && (u == thr);
    const unsigned int ballot = __ballot_sync(0xffffffffu, tie);
    if (lane == 0) warp_cnt[warp] = __popc(ballot);
    __syncthreads();
    unsigned long long before = taken_s;
    for (int wdx = 0; wdx < warp; ++wdx) before += warp_cnt[wdx];
    before += __popc(ballot & ((1u << lane) - 1u));
    if (i < n) {
      if (u < thr) mask[i] = 0;
      else if (tie && before < n_ties) mask[i] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t = taken_s;
      for (int wdx = 0; wdx < 32; ++wdx) t += warp_cnt[wdx];
      taken_s = t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *tie_counter = taken_s;
}

// out[r] = sum_c |w[r, c]|  -- one warp per row, fp64 accumulation
__global__ void row_abs_sums_kernel(const float* __restrict__ w, long long ld, double* __restrict__ out, int rows,
                                    int cols) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  double s = 0.0;
  for (int c = lane; c < cols; c += 32) s += fabs(static_cast<double>(w[static_cast<long long>(row) * ld + c]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}
// out[c] = sum_r |w[r, c]|  -- one thread per column (coalesced across the warp)
__global__ void col_abs_sums_kernel(const float* __restrict__ w, long long ld, double* __restrict__ out, int rows,
                                    int cols) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double s = 0.0;
  for (int r = 0; r < rows; ++r) s += fabs(static_cast<double>(w[static_cast<long long>(r) * ld + c]));
  out[c] = s;
}
}  // namespace mh

using namespace mh;
#define ST reinterpret_cast<cudaStream_t>(stream)

extern "C" int mh_abs_kth_smallest(const float* const* ptrs, const long long* sizes, int n_tensors, long long k,
                                   unsigned long long* workspace, unsigned long long* result, void* stream) {
  MH_CHECK(k >= 1, "kth_smallest: k must be >= 1");
  long long total = 0;
  for (int t = 0; t < n_tensors; ++t) total += sizes[t];
  MH_CHECK(k <= total, "kth_smallest: k (%lld) exceeds the number of elements (%lld)", k, total);
  radix_init_kernel<<<1, 256, 0, ST>>>(workspace, static_cast<unsigned long long>(k));
  MH_LAUNCH_CHECK();
  ++g_launches;
  const int shifts[3] = {20, 9, 0}, bits[3] = {11, 11, 9};  // bit 31 (sign) is cleared: 31 bits total
  for (int pass = 0; pass < 3; ++pass) {
    for (int t = 0; t < n_tensors; ++t) {
      if (sizes[t] == 0) continue;
      long long g = (sizes[t] + 512 * 8 - 1) / (512 * 8);
      const long long cap = static_cast<long long>(sm_count()) * 4;
      if (g > cap) g = cap;
      radix_hist_kernel<<<static_cast<int>(g), 512, 0, ST>>>(ptrs[t], sizes[t], shifts[pass], bits[pass], workspace);
      MH_LAUNCH_CHECK();
      ++g_launches;
    }
    radix_pick_kernel<<<1, 1, 0, ST>>>(workspace, shifts[pass], bits[pass], result, pass == 2);
    MH_LAUNCH_CHECK();
    ++g_launches;
  }
  return 0;
}

extern "C" int mh_apply_threshold_mask(const float* w, uint8_t* mask, long long n, const unsigned long long* result,
                                       long long k, unsigned long long* tie_counter, void* stream) {
  if (n == 0) return 0;
  apply_threshold_kernel<<<1, 1024, 0, ST>>>(w, mask, n, result, static_cast<unsigned long long>(k), tie_counter);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

extern "C" int mh_row_abs_sums(const float* w, long long ld, double* out, int rows, int cols, void* stream) {
  row_abs_sums_kernel<<<(rows + 7) / 8, 256, 0, ST>>>(w, ld, out, rows, cols);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}
extern "C" int mh_col_abs_sums(const float* w, long long ld, double* out, int rows, int cols, void* stream) {
  col_abs_sums_kernel<<<(cols + 127) / 128, 128, 0, ST>>>(w, ld, out, rows, cols);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}
