// LayerNorm forward / backward and bias-gradient column sums: fp32 statistics, warp-shuffle reductions, packed fp32
// arithmetic.  Rows reach a warp through its private 3-deep cp.async ring in shared memory (16-byte chunks, lane-private
// slots): two rows are in flight while one is reduced and no registers are spent on prefetch.  Forward: one warp per row,
// gamma / beta resident in registers.  Backward: NCH warps per row (one 8-column chunk per lane), see ln_bwd_kernel.
#include "mh_b200.h"
#define MH_PDL_FAMILY 4
#include "mh_common.cuh"
#include "mh_ptx.cuh"

namespace mh {
extern long long g_launches;

constexpr int LN_WARPS = 8;

constexpr int LNB_STAGES = 3;
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// (ld.volatile: ptxas otherwise merges the second pass's reads with the first pass's and keeps the whole row in
//  registers -- exactly the pressure the shared-memory ring is there to remove)
__device__ __forceinline__ uint4 lds128(uint32_t smem_addr) {
  uint4 v;
  asm volatile("ld.volatile.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_addr) : "memory");
  return v;
}
constexpr int ln_fwd_smem_bytes(int nch) { return LN_WARPS * LNB_STAGES * nch * 512; }

// NCH = number of 8-element chunks per lane (cols <= NCH * 256); EXACT: cols == NCH * 256, no chunk predicates
// (the encoder's 768 and 512 columns).
// ncu on the first ring version (profiles/r01_q_ncu_ln_fwd.txt, r02 re-capture): L1/TEX throughput 78 %, DRAM 22 % -- 6 of the
// 10.5 KB a row moved through L1 were gamma / beta reloads.  With them resident the kernel streams at 6 TB/s (12.4 us per
// launch in the step against 18.1).
#ifndef LN_FWD_MINB
#define LN_FWD_MINB 3
#endif
template <int NCH, bool EXACT, bool HAS_DROP>
__global__ void __launch_bounds__(LN_WARPS * 32, LN_FWD_MINB)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows,
              int cols_rt, float eps, const DropCfg drop) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t lnf_ring[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warp_global = blockIdx.x * LN_WARPS + warp;
  const int nwarps = gridDim.x * LN_WARPS;
  const int cols = EXACT ? NCH * 256 : cols_rt;
  const int nchunks = EXACT ? NCH * 32 : cols_rt >> 3;
  const float inv_cols = 1.0f / static_cast<float>(cols);
  const DropState dstate(drop);
  const uint32_t ring = static_cast<uint32_t>(__cvta_generic_to_shared(lnf_ring)) + warp * (LNB_STAGES * NCH * 512) + lane * 16;
  auto issue = [&](int row, int stage) {
    if (row < rows) {
      const __nv_bfloat16* xr = x + static_cast<long long>(row) * cols + lane * 8;
#pragma unroll
      for (int i = 0; i < NCH; ++i)
        if (EXACT || lane + 32 * i < nchunks) cp_async16(ring + (stage * NCH + i) * 512, xr + i * 256);
    }
    cp_async_commit();
  };
  issue(warp_global, 0);
  issue(warp_global + nwarps, 1);
  // gamma / beta of this lane's columns live in registers for the whole kernel (48 registers at 768 columns): reloading
  // them for every row put 6 KB per row through L1 next to 4.5 KB of row traffic (ring fill, ring read, store) and made the
  // L1 / shared-memory data path, not HBM, the limit (20.7 us against 8.8 us for a copy of the same tensor)
  uint64_t g[NCH][4], bt[NCH][4];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    if (EXACT || lane + 32 * i < nchunks) {
      const int c8 = lane * 8 + i * 256;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c8 + 4));
      g[i][0] = pack2f(g0.x, g0.y); g[i][1] = pack2f(g0.z, g0.w); g[i][2] = pack2f(g1.x, g1.y); g[i][3] = pack2f(g1.z, g1.w);
      bt[i][0] = pack2f(b0.x, b0.y); bt[i][1] = pack2f(b0.z, b0.w); bt[i][2] = pack2f(b1.x, b1.y); bt[i][3] = pack2f(b1.z, b1.w);
    }
  }
  int stage = 0;
  for (int row = warp_global; row < rows; row += nwarps) {
    issue(row + 2 * nwarps, stage + 2 >= LNB_STAGES ? stage + 2 - LNB_STAGES : stage + 2);
    // keep bits of this lane's chunks (bit-sliced dropout: 32 decisions per generated word), produced while the row is
    // still in flight -- the kernel is latency-bound, the Philox rounds are free here
    uint32_t kb[NCH];
    if (HAS_DROP) dstate.keep_bytes_row<NCH>(drop, static_cast<uint64_t>(row) * (nchunks >> 2), nchunks, lane, kb);
    cp_async_wait<2>();
    // packed fp32 pairs (FADD2 / FFMA2): the kernel is issue- and latency-bound (a plain copy of the same L2-resident
    // tensor takes 8.8 us), not HBM-bound -- 10.9 executed instructions per element before, ~7 now
    // the row stays in registers as raw bf16 (12 registers at 768 columns) and is unpacked once per pass: with gamma / beta
    // resident, fp32 copies of the row would push the kernel below three blocks per SM
    uint4 raw[NCH];
    auto unpack = [](const uint4& q, uint64_t (&o)[4]) {
      o[0] = pack2f(bf16_lo(q.x), bf16_hi(q.x));
      o[1] = pack2f(bf16_lo(q.y), bf16_hi(q.y));
      o[2] = pack2f(bf16_lo(q.z), bf16_hi(q.z));
      o[3] = pack2f(bf16_lo(q.w), bf16_hi(q.w));
    };
    uint64_t sa = pack2f(0.f, 0.f), sb = sa;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (EXACT || lane + 32 * i < nchunks) {
        raw[i] = lds128(ring + (stage * NCH + i) * 512);
        uint64_t v[4];
        unpack(raw[i], v);
        sa = fadd2(sa, fadd2(v[0], v[1]));  // two accumulators: half the dependent chain
        sb = fadd2(sb, fadd2(v[2], v[3]));
      }
    }
    float s0, s1;
    unpack2f(fadd2(sa, sb), s0, s1);
    const float mean = warp_sum(s0 + s1) * inv_cols;
    const uint64_t nmean2 = pack2f(-mean, -mean);
    uint64_t qa = pack2f(0.f, 0.f), qb = qa;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (EXACT || lane + 32 * i < nchunks) {
        uint64_t v[4];
        unpack(raw[i], v);
#pragma unroll
        for (int k = 0; k < 4; k += 2) {
          const uint64_t d0 = fadd2(v[k], nmean2), d1 = fadd2(v[k + 1], nmean2);
          qa = ffma2(d0, d0, qa);
          qb = ffma2(d1, d1, qb);
        }
      }
    }
    unpack2f(fadd2(qa, qb), s0, s1);
    const float rstd = rsqrtf(warp_sum(s0 + s1) * inv_cols + eps);
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
    __nv_bfloat16* yr = y + static_cast<long long>(row) * cols + lane * 8;
    const float nmr = -mean * rstd;
    const uint64_t rstd2 = pack2f(rstd, rstd), nmr2 = pack2f(nmr, nmr);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (EXACT || lane + 32 * i < nchunks) {
        uint64_t v[4];
        unpack(raw[i], v);
        float o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) unpack2f(ffma2(ffma2(v[k], rstd2, nmr2), g[i][k], bt[i][k]), o[2 * k], o[2 * k + 1]);
        if (HAS_DROP) DropState::apply8(drop, kb[i], o);
        stg128(yr + i * 256, f32_to_bf16x8(o));
      }
    }
    stage = stage + 1 == LNB_STAGES ? 0 : stage + 1;
  }
  cp_async_wait<0>();
}

// Backward.  dy_eff = dy (* keep mask of the output dropout `din`, if the forward dropped the
// LN output).  dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)).
// Also emits dx_drop = dx * keep mask of `dout` (the dropout that sat on the GEMM output
// feeding this LayerNorm's input) so the dgrad/wgrad GEMMs can consume it directly, and (HAS_DCOL) the column sums of
// that second output -- the bias gradient of the GEMM that fed this LayerNorm's input (out_proj / fc2), which otherwise
// costs a separate pass over the same tensor.
//
// History: (1) one warp per row, row and prefetched successor in registers: 42 instructions per element, issue-bound,
// 3.1 TB/s.  (2) rows through a per-warp cp.async ring, packed fp32: 36 us for the step's form, 16 warps per SM at 124
// registers (72 of them the per-lane dgamma / dbeta / dcol accumulators of 24 columns) and, per ncu, the L1 / shared-memory
// data path as the busiest unit (67 %): 18 KB per row went through it (ring fill 3, two passes of ring reads 6, gamma
// twice 6, stores 3).  (3) now: NCH warps per row, one 8-column chunk per lane.  gamma (8 registers), the raw bf16 row
// (8) and the accumulators (24) all stay in registers, so a row costs 9 KB of L1 traffic (ring fill, one ring read,
// stores) and the kernel runs 24+ warps per SM; the row statistics cross the NCH warps through 8 bytes of shared memory
// and one named barrier per row; the dropout words of four consecutive rows of a warp come from ONE Philox pass (lane
// l generates word l & 7 of row l >> 3).
#ifndef LN_BWD_MINB
#define LN_BWD_MINB 2
#endif
constexpr int LNB_RPB = 4;  // row slots per block: a block is LNB_RPB x NCH warps
constexpr int ln_bwd_smem_bytes(int nch) { return LNB_RPB * nch * LNB_STAGES * 2 * 512; }

template <int NCH, bool EXACT, bool HAS_DIN, bool HAS_DCOL>
__global__ void __launch_bounds__(LNB_RPB * NCH * 32, LN_BWD_MINB)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
              const float* __restrict__ gamma, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dx_drop, float* __restrict__ dgamma,
              float* __restrict__ dbeta, float* __restrict__ dcol, int rows, int cols_rt, const DropCfg din, const DropCfg dout) {
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t lnb_ring[];
  __shared__ float2 part_stats[LNB_RPB][2][NCH];            // (s1, s2) partial sums of a row, double-buffered by row parity
  __shared__ __align__(16) float red[LNB_RPB][NCH * 256 + 4];  // final cross-slot reduction of the column accumulators
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int slot = warp / NCH, part = warp - slot * NCH;
  const int row_stride = gridDim.x * LNB_RPB;
  const int cols = EXACT ? NCH * 256 : cols_rt;
  const int nchunks = EXACT ? NCH * 32 : cols_rt >> 3;
  const int chunk = part * 32 + lane;  // this lane's 8-column chunk of the row
  const bool has_chunk = EXACT || chunk < nchunks;
  const float inv_cols = 1.0f / static_cast<float>(cols);
  const bool drop_out = dx_drop != nullptr && dout.thresh != 0;
  const DropState st_in(din), st_out(dout);  // (st_in is dead code unless HAS_DIN)
  // ring slot of (stage, dy | x) for this lane: 512 bytes per (stage, tensor), 16 bytes per lane
  const uint32_t ring = static_cast<uint32_t>(__cvta_generic_to_shared(lnb_ring)) + warp * (LNB_STAGES * 2 * 512) + lane * 16;
  auto slot_addr = [&](int stage, int which) -> uint32_t { return ring + (stage * 2 + which) * 512; };
  auto issue = [&](int row, int stage) {
    if (row < rows && has_chunk) {
      const long long off = static_cast<long long>(row) * cols + chunk * 8;
      cp_async16(slot_addr(stage, 0), dy + off);
      cp_async16(slot_addr(stage, 1), x + off);
    }
    cp_async_commit();  // (an empty group past the end keeps the wait_group count uniform)
  };
  auto unpack8 = [](uint4 q, uint64_t (&o)[4]) {
    o[0] = pack2f(bf16_lo(q.x), bf16_hi(q.x));
    o[1] = pack2f(bf16_lo(q.y), bf16_hi(q.y));
    o[2] = pack2f(bf16_lo(q.z), bf16_hi(q.z));
    o[3] = pack2f(bf16_lo(q.w), bf16_hi(q.w));
  };
  // the keep words of FOUR consecutive rows of this warp come from one Philox pass: lane l generates word (l & 7) of this
  // warp's 8 words (32 chunks x 8 bits) of row `row + (l >> 3) * row_stride`; every row then fetches its byte by shuffle
  auto gen_keep4 = [&](const DropState& st, const DropCfg& d, int row) -> uint32_t {
    const int r = row + (lane >> 3) * row_stride;
    const int w = part * 8 + (lane & 7);
    return (r < rows && w < (nchunks >> 2)) ? st.keep32(d, static_cast<uint64_t>(r) * (nchunks >> 2) + w) : 0xffffffffu;
  };
  auto keep_byte = [&](uint32_t cache, int sub) -> uint32_t {
    return (__shfl_sync(0xffffffffu, cache, sub * 8 + (lane >> 2)) >> (8 * (lane & 3))) & 0xffu;
  };

  const uint64_t zero2 = pack2f(0.f, 0.f);
  uint64_t dg[4], db[4], dc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) dg[k] = db[k] = dc[k] = zero2;
  uint64_t g[4] = {zero2, zero2, zero2, zero2};
  if (has_chunk) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + chunk * 8));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + chunk * 8 + 4));
    g[0] = pack2f(g0.x, g0.y); g[1] = pack2f(g0.z, g0.w); g[2] = pack2f(g1.x, g1.y); g[3] = pack2f(g1.z, g1.w);
  }

  const int row0 = blockIdx.x * LNB_RPB + slot;
  issue(row0, 0);
  issue(row0 + row_stride, 1);
  float nmean = 0.f, nrstd = 0.f;
  if (row0 < rows) {
    nmean = __ldg(mean_in + row0);
    nrstd = __ldg(rstd_in + row0);
  }
  int stage = 0, it = 0;
  uint32_t kin_cache = 0xffffffffu, kout_cache = 0xffffffffu;
  // every warp of a row slot walks the same rows: the named barrier below is uniform within the slot
  for (int row = row0; row < rows; row += row_stride, ++it) {
    {  // refill the ring slot that was consumed in the previous trip (this thread's reads of it have retired)
      const int s2 = stage + 2 >= LNB_STAGES ? stage + 2 - LNB_STAGES : stage + 2;
      issue(row + 2 * row_stride, s2);
    }
    const float rstd = nrstd, nmr = -nmean * nrstd;
    if (row + row_stride < rows) {
      nmean = __ldg(mean_in + row + row_stride);
      nrstd = __ldg(rstd_in + row + row_stride);
    }
    if ((it & 3) == 0) {  // (generated under the loads)
      if (HAS_DIN) kin_cache = gen_keep4(st_in, din, row);
      if (drop_out) kout_cache = gen_keep4(st_out, dout, row);
    }
    cp_async_wait<2>();  // everything but the two youngest groups has landed: this row is in its slot
    uint64_t d[4] = {zero2, zero2, zero2, zero2}, xv[4] = {zero2, zero2, zero2, zero2};
    uint32_t kin_b = 0xffu, kout_b = 0xffu;  // (the shuffles need all 32 lanes: outside the per-chunk predicate)
    if (HAS_DIN) kin_b = keep_byte(kin_cache, it & 3);
    if (drop_out) kout_b = keep_byte(kout_cache, it & 3);
    if (has_chunk) {
      const uint4 qd = lds128(slot_addr(stage, 0));
      if (HAS_DIN) {
        float f[8];
        bf16x8_to_f32(qd, f);
        DropState::apply8(din, kin_b, f);
#pragma unroll
        for (int k = 0; k < 4; ++k) d[k] = pack2f(f[2 * k], f[2 * k + 1]);
      } else {
        unpack8(qd, d);
      }
      unpack8(lds128(slot_addr(stage, 1)), xv);
    }
    // pass 1: xhat, the two row sums, the column accumulators
    uint64_t s1p = zero2, s2p = zero2, xh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      xh[k] = ffma2(xv[k], pack2f(rstd, rstd), pack2f(nmr, nmr));
      const uint64_t t = fmul2(d[k], xh[k]);  // dy * xhat
      s1p = ffma2(d[k], g[k], s1p);
      s2p = ffma2(t, g[k], s2p);
      dg[k] = fadd2(dg[k], t);
      db[k] = fadd2(db[k], d[k]);
    }
    // (lanes past the last chunk hold dy = 0: they add nothing to the sums and store nothing)
    float s1, s2;
    {
      float a0, a1, b0, b1;
      unpack2f(s1p, a0, a1);
      unpack2f(s2p, b0, b1);
      s1 = a0 + a1;
      s2 = b0 + b1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {  // the two reductions interleaved: one latency chain instead of two
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (NCH > 1) {  // across the row's NCH warps: 8 bytes each through shared memory, one named barrier
        if (lane == 0) part_stats[slot][it & 1][part] = make_float2(s1, s2);
        bar_sync(1 + slot, NCH * 32);
        s1 = 0.f; s2 = 0.f;
#pragma unroll
        for (int w = 0; w < NCH; ++w) {
          const float2 ps = part_stats[slot][it & 1][w];
          s1 += ps.x;
          s2 += ps.y;
        }
      }
      s1 *= inv_cols;
      s2 *= inv_cols;
    }
    // pass 2: dx = rstd (d g - s1 - xh s2)
    const float c0 = -rstd * s1, c1 = -rstd * s2;
    const uint64_t c02 = pack2f(c0, c0), c12 = pack2f(c1, c1), rstd2 = pack2f(rstd, rstd);
    float o[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) unpack2f(ffma2(d[k], fmul2(g[k], rstd2), ffma2(xh[k], c12, c02)), o[2 * k], o[2 * k + 1]);
    if (has_chunk) {
      const long long off = static_cast<long long>(row) * cols + chunk * 8;
      stg128(dx + off, f32_to_bf16x8(o));
      if (dx_drop != nullptr) {
        if (drop_out) DropState::apply8(dout, kout_b, o);
        stg128(dx_drop + off, f32_to_bf16x8(o));
      }
      if (HAS_DCOL) {
#pragma unroll
        for (int k = 0; k < 4; ++k) dc[k] = fadd2(dc[k], pack2f(o[2 * k], o[2 * k + 1]));
      }
    }
    stage = stage + 1 == LNB_STAGES ? 0 : stage + 1;
  }
  cp_async_wait<0>();
  // block reduction of the per-warp partial dgamma / dbeta (/ dcol) over the row slots, then one atomic per column
  for (int pass = 0; pass < (HAS_DCOL ? 3 : 2); ++pass) {
    __syncthreads();
    {
      float f[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) unpack2f(pass == 0 ? dg[k] : (pass == 1 ? db[k] : dc[k]), f[2 * k], f[2 * k + 1]);
      *reinterpret_cast<float4*>(&red[slot][chunk * 8]) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(&red[slot][chunk * 8 + 4]) = make_float4(f[4], f[5], f[6], f[7]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cols; c += LNB_RPB * NCH * 32) {
      float t = 0.f;
#pragma unroll
      for (int sidx = 0; sidx < LNB_RPB; ++sidx) t += red[sidx][c];
      atomicAdd((pass == 0 ? dgamma : (pass == 1 ? dbeta : dcol)) + c, t);
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

template <int NCH, bool EXACT, bool HAS_DROP, typename... Args>
static int launch_ln_fwd(int grid, cudaStream_t st, Args... args) {
  auto kfn = ln_fwd_kernel<NCH, EXACT, HAS_DROP>;
  static bool configured = false;
  if (!configured) {
    MH_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, ln_fwd_smem_bytes(NCH)));
    configured = true;
  }
  MH_CUDA(launch_pdl(kfn, dim3(grid), dim3(LN_WARPS * 32), ln_fwd_smem_bytes(NCH), st, args...));
  return 0;
}

template <int NCH, bool EXACT, bool HAS_DIN, bool HAS_DCOL, typename... Args>
static int launch_ln_bwd(int grid, cudaStream_t st, Args... args) {
  auto kfn = ln_bwd_kernel<NCH, EXACT, HAS_DIN, HAS_DCOL>;
  static bool configured = false;  // the row ring needs the > 48 KB opt-in
  if (!configured) {
    MH_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, ln_bwd_smem_bytes(NCH)));
    configured = true;
  }
  MH_CUDA(launch_pdl(kfn, dim3(grid), dim3(LNB_RPB * NCH * 32), ln_bwd_smem_bytes(NCH), st, args...));
  return 0;
}

// Narrow matrices (cols <= 1024: the [B*T, 768] gradients behind the out_proj / fc2 bias gradients): the slab kernel
// above has 3 column slabs x ~200 row slices of ~120 rows -- too little work per block, 2.6 TB/s.  Here a warp owns
// whole rows (NCH 16-byte chunks per lane) and keeps FOUR rows in flight per trip; per-lane column sums stay in
// registers and leave through the same block reduction + one atomic per column.
template <int NCH>
__global__ void __launch_bounds__(256)
colsum_rows_kernel(const __nv_bfloat16* __restrict__ x, long long ld, float* __restrict__ out, int rows, int cols) {
  pdl_prologue();
  __shared__ float red[8][32 * 8 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = gridDim.x * 8;
  const int nchunks = cols >> 3;
  float acc[NCH][8];
#pragma unroll
  for (int i = 0; i < NCH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  int row = blockIdx.x * 8 + warp;
  for (; row + 3 * nwarps < rows; row += 4 * nwarps) {
    uint4 q[4][NCH];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < NCH; ++i)
        if (lane + 32 * i < nchunks) q[u][i] = ldg128(x + static_cast<long long>(row + u * nwarps) * ld + lane * 8 + i * 256);
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < NCH; ++i)
        if (lane + 32 * i < nchunks) {
          float v[8];
          bf16x8_to_f32(q[u][i], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] += v[j];
        }
  }
  for (; row < rows; row += nwarps) {
#pragma unroll
    for (int i = 0; i < NCH; ++i)
      if (lane + 32 * i < nchunks) {
        float v[8];
        bf16x8_to_f32(ldg128(x + static_cast<long long>(row) * ld + lane * 8 + i * 256), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] += v[j];
      }
  }
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[i][j];
    __syncthreads();
    const int col = i * 256 + threadIdx.x;
    if (col < cols) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
      atomicAdd(out + col, t);
    }
  }
}

template <typename F>
static int dispatch_nch(int cols, F&& f) {
  const int nch = (cols / 8 + 31) / 32;
  if (cols == nch * 256) {  // whole chunks only: predicate-free kernels
    switch (nch) {
      case 1: return f(std::integral_constant<int, 1>(), std::true_type());
      case 2: return f(std::integral_constant<int, 2>(), std::true_type());
      case 3: return f(std::integral_constant<int, 3>(), std::true_type());
      case 4: return f(std::integral_constant<int, 4>(), std::true_type());
      default: break;
    }
  }
  switch (nch) {
    case 1: return f(std::integral_constant<int, 1>(), std::false_type());
    case 2: return f(std::integral_constant<int, 2>(), std::false_type());
    case 3: return f(std::integral_constant<int, 3>(), std::false_type());
    case 4: return f(std::integral_constant<int, 4>(), std::false_type());
    default: set_error("LayerNorm supports up to 1024 columns (got %d)", cols); return 1;
  }
}
}  // namespace mh

using namespace mh;

extern "C" int mh_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                float* rstd, int rows, int cols, float eps, float p_drop, uint64_t seed,
                                uint32_t site, void* stream) {
  MH_CHECK(rows > 0 && cols > 0 && cols % 8 == 0, "layernorm: bad shape %d x %d", rows, cols);
  MH_CHECK(!(p_drop > 0.f) || cols % 32 == 0, "layernorm: dropout needs cols %% 32 == 0 (one dropout stream word = 32 elements), got %d", cols);
  MH_CHECK(x != nullptr && y != nullptr && gamma != nullptr && beta != nullptr && mean != nullptr && rstd != nullptr,
           "layernorm: null pointer");
  MH_CHECK(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
             reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
           "layernorm: x, y, gamma and beta must be 16-byte aligned (rows are moved as 16-byte chunks)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // resident blocks only (LN_FWD_MINB per SM: gamma / beta stay in registers): a warp then walks ~10 rows at the bench shape
  // and its ring stays primed
  const int grid = min((rows + LN_WARPS - 1) / LN_WARPS, sm_count() * LN_FWD_MINB);
  const DropCfg d = make_drop(p_drop, seed, site);
  return dispatch_nch(cols, [&](auto nch, auto exact) {
    const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
    __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
    const int rc = d.thresh != 0
        ? launch_ln_fwd<decltype(nch)::value, decltype(exact)::value, true>(grid, st, xp, gamma, beta, yp, mean, rstd, rows, cols, eps, d)
        : launch_ln_fwd<decltype(nch)::value, decltype(exact)::value, false>(grid, st, xp, gamma, beta, yp, mean, rstd, rows, cols, eps, d);
    if (rc != 0) return rc;
    ++g_launches;
    return 0;
  });
}

static int ln_bwd_impl(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd, void* dx,
                       void* dx_drop, float* dgamma, float* dbeta, float* dcol, int rows, int cols, float p_in,
                       uint64_t seed_in, uint32_t site_in, float p_out, uint64_t seed_out, uint32_t site_out, void* stream) {
  MH_CHECK(rows > 0 && cols > 0 && cols % 8 == 0, "layernorm_bwd: bad shape %d x %d", rows, cols);
  MH_CHECK(!(p_in > 0.f || p_out > 0.f) || cols % 32 == 0, "layernorm_bwd: dropout needs cols %% 32 == 0, got %d", cols);
  MH_CHECK(dy != nullptr && x != nullptr && gamma != nullptr && mean != nullptr && rstd != nullptr && dx != nullptr &&
               dgamma != nullptr && dbeta != nullptr,
           "layernorm_bwd: null pointer");
  MH_CHECK(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) |
             reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dx_drop)) & 15) == 0,
           "layernorm_bwd: dy, x, gamma, dx and dx_drop must be 16-byte aligned (rows are moved as 16-byte chunks)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = min((rows + LNB_RPB - 1) / LNB_RPB, sm_count() * LN_BWD_MINB);
  const DropCfg din = make_drop(p_in, seed_in, site_in), dout = make_drop(p_out, seed_out, site_out);
  const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(dy);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* dxp = reinterpret_cast<__nv_bfloat16*>(dx);
  __nv_bfloat16* ddp = reinterpret_cast<__nv_bfloat16*>(dx_drop);
  return dispatch_nch(cols, [&](auto nch, auto exact) {
    constexpr int N = decltype(nch)::value;
    constexpr bool E = decltype(exact)::value;
    int rc;
    if (dcol != nullptr)
      rc = din.thresh != 0
               ? launch_ln_bwd<N, E, true, true>(grid, st, dyp, xp, gamma, mean, rstd, dxp, ddp, dgamma, dbeta, dcol, rows, cols, din, dout)
               : launch_ln_bwd<N, E, false, true>(grid, st, dyp, xp, gamma, mean, rstd, dxp, ddp, dgamma, dbeta, dcol, rows, cols, din, dout);
    else
      rc = din.thresh != 0
               ? launch_ln_bwd<N, E, true, false>(grid, st, dyp, xp, gamma, mean, rstd, dxp, ddp, dgamma, dbeta, dcol, rows, cols, din, dout)
               : launch_ln_bwd<N, E, false, false>(grid, st, dyp, xp, gamma, mean, rstd, dxp, ddp, dgamma, dbeta, dcol, rows, cols, din, dout);
    if (rc != 0) return rc;
    ++g_launches;
    return 0;
  });
}

extern "C" int mh_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                                const float* rstd, void* dx, void* dx_drop, float* dgamma, float* dbeta, int rows,
                                int cols, float p_in, uint64_t seed_in, uint32_t site_in, float p_out,
                                uint64_t seed_out, uint32_t site_out, void* stream) {
  return ln_bwd_impl(dy, x, gamma, mean, rstd, dx, dx_drop, dgamma, dbeta, nullptr, rows, cols, p_in, seed_in, site_in, p_out,
                     seed_out, site_out, stream);
}

// same, and dcol[c] += sum over rows of the output (dx_drop when given, else dx): the bias gradient of the linear layer
// whose output fed this LayerNorm (module.py:121-123 / 129-131)
extern "C" int mh_layernorm_bwd_colsum(const void* dy, const void* x, const float* gamma, const float* mean,
                                       const float* rstd, void* dx, void* dx_drop, float* dgamma, float* dbeta, float* dcol,
                                       int rows, int cols, float p_in, uint64_t seed_in, uint32_t site_in, float p_out,
                                       uint64_t seed_out, uint32_t site_out, void* stream) {
  MH_CHECK(dcol != nullptr, "layernorm_bwd_colsum: null dcol");
  return ln_bwd_impl(dy, x, gamma, mean, rstd, dx, dx_drop, dgamma, dbeta, dcol, rows, cols, p_in, seed_in, site_in, p_out,
                     seed_out, site_out, stream);
}

extern "C" int mh_colsum(const void* x, long long ld, float* out, int rows, int cols, void* stream) {
  MH_CHECK(rows > 0 && cols > 0 && cols % 8 == 0 && ld % 8 == 0 && ld >= cols, "colsum: bad shape %d x %d (ld %lld)", rows, cols, ld);
  MH_CHECK(x != nullptr && out != nullptr && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
           "colsum: x must be a 16-byte aligned bf16 matrix");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cols <= 1024 && rows >= 4096) {
    const int grid = sm_count() * 2;
    const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
    switch ((cols + 255) / 256) {
      case 1: MH_CUDA(launch_pdl(colsum_rows_kernel<1>, dim3(grid), dim3(256), 0, st, xp, ld, out, rows, cols)); break;
      case 2: MH_CUDA(launch_pdl(colsum_rows_kernel<2>, dim3(grid), dim3(256), 0, st, xp, ld, out, rows, cols)); break;
      case 3: MH_CUDA(launch_pdl(colsum_rows_kernel<3>, dim3(grid), dim3(256), 0, st, xp, ld, out, rows, cols)); break;
      default: MH_CUDA(launch_pdl(colsum_rows_kernel<4>, dim3(grid), dim3(256), 0, st, xp, ld, out, rows, cols)); break;
    }
    ++g_launches;
    return 0;
  }
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
