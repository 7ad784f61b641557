// Positional grouped convolution of the MelHuBERT encoder on tcgen05 (reference module.py:175-188,229-231:
// weight-normed Conv1d(768, 768, k = 128, padding = 64, groups = 16) -> SamePad -> GELU, added to its input).
//
// Per group g the convolution is a GEMM  Y_g[B*T, 48] = Xwin_g[B*T, 128 taps x 48 ch] * W_g[48, 128 x 48]^T
// whose A operand is a sliding window.  The window is never materialised: the CTA keeps the input rows
// of its time slab ONCE in shared memory, in a "chunk-column" layout (for each 8-channel chunk, the rows
// follow each other at a 16-byte pitch).  In the un-swizzled UMMA canonical layout that makes a row
// shift a pure start-address change (+16 B per row), so the A descriptor of tap k is the descriptor of tap 0
// advanced by k rows.  The same layout read MN-major gives the weight-gradient kernel its Hankel operand
// (M index = (tap j, channel c) -> address base + 16 j + 2 c: overlapping 16-byte chunks).
//
//   posconv_kernel      : forward (z = conv + bias saved, y = x + gelu(z)) and input gradient
//                         (dx = dy + conv^T(dz): same kernel, taps flipped / in-out swapped in the weights)
//   posconv_wgrad_kernel: dW[co, ci, k] += sum_t dz[t, co] * x[t + k - 64, ci]
//   weight prep / weight-norm backward / GELU-backward elementwise helpers
#include "mh_b200.h"
#include "mh_common.cuh"
#include "mh_ptx.cuh"

namespace mh {
extern long long g_launches;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

constexpr int PC_CG = 48;     // channels per group (in and out)
constexpr int PC_TAPS = 128;
constexpr int PC_CHUNKS = PC_CG / 8;          // 16-byte channel chunks per group
constexpr int PC_TAP_BYTES = PC_CG * PC_CG * 2;  // one tap of one group: [6 chunks][48 rows][8] bf16 = 4608 B

// un-swizzled UMMA shared-memory descriptor: lbo / sbo in bytes (see mh_ptx.cuh for the field map)
__device__ __forceinline__ uint64_t make_sdesc_ns(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// forward / dgrad
// ------------------------------------------------------------------------------------------------
// Two taps per MMA.  With one tap per MMA (N = 48) every 128 x 48 x 16 instruction reads 5.5 KB of operands for 24 clk of
// math: 229 B/clk against the 128 B/clk of shared memory, i.e. ~55 % of the tensor peak by construction.  The A rows of
// tap 2m (window rows base + 2m + r) multiplied by B = [W_2m | W_2m+1] (N = 96) give D0[r] (the tap-2m term of output row
// r) and D1[r] = x[base + 2m + r] W_2m+1 -- the tap-(2m+1) term of output row r - 1.  So out[r] = D0[r] + D1[r + 1]:
// the second accumulator is read one TMEM lane up in the epilogue, a 128-row tile yields 127 output rows, and the A
// operand traffic per FLOP halves.
constexpr int PCF_NT = 2;                        // 128-row MMA tiles per CTA
constexpr int PCF_TSTRIDE = 127;                 // output rows per tile (see above)
constexpr int PCF_ROWS = 384;                    // window rows kept in smem (127 + 126 + 127 + 1 = 381 used)
constexpr int PCF_CS = PCF_ROWS * 16;            // chunk stride in bytes
constexpr int PCF_TPS = 4;                       // taps per weight stage (= 2 tap pairs)
constexpr int PCF_STAGES = 3;
constexpr int PC_PAIR_BYTES = 2 * PC_TAP_BYTES;  // one tap pair of one group: [6 chunks][96 rows = 2 taps x 48 co][8] bf16
constexpr int PCF_STAGE_BYTES = PCF_TPS * PC_TAP_BYTES;
constexpr int PCF_XCH_BYTES = 2 * 4 * PC_CG * 4; // D1 rows of the warps' first lanes (lane 31 needs the next warp's lane 0), x2 tiles
constexpr int PCF_SMEM = PC_CHUNKS * PCF_CS + PCF_STAGES * PCF_STAGE_BYTES + PCF_XCH_BYTES + 128 + 128 /*align*/;
constexpr int PCF_THREADS = 192;                 // warps 0-3 epilogue, 4 = loads, 5 = MMA
constexpr int PCF_TMEM_TILE = 96;                // accumulator columns per tile: D0 | D1

struct PosConvParams {
  const __nv_bfloat16* w;      // [G][64 tap pairs][6 chunks][96 = (tap & 1) * 48 + co][8]
  const float* bias;           // [C] or null
  const __nv_bfloat16* res;    // [B*T, C] residual added in the epilogue
  __nv_bfloat16* z;            // [B*T, C] pre-activation output (mode 0) or null
  __nv_bfloat16* y;            // [B*T, C]
  int B, T, C, pad_left, mode; // mode 0: y = res + gelu(acc + bias), z = acc + bias;  mode 1: y = res + acc
};

__global__ void __launch_bounds__(PCF_THREADS, 2)
posconv_kernel(const __grid_constant__ CUtensorMap tmX, const PosConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* sX = smem;
  uint8_t* sW = sX + PC_CHUNKS * PCF_CS;
  float* sXch = reinterpret_cast<float*>(sW + PCF_STAGES * PCF_STAGE_BYTES);  // [tile][warp][48]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sXch + PCF_XCH_BYTES / 4);
  uint64_t* x_full = bars;
  uint64_t* w_full = bars + 1;                 // [PCF_STAGES]
  uint64_t* w_empty = w_full + PCF_STAGES;     // [PCF_STAGES]
  uint64_t* acc_full = w_empty + PCF_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slab = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
  const int t0 = slab * (PCF_NT * PCF_TSTRIDE);
  const int n_tiles = min(PCF_NT, (p.T - t0 + PCF_TSTRIDE - 1) / PCF_TSTRIDE);

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmX);
    mbar_init(x_full, 1);
    for (int s = 0; s < PCF_STAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (elect_one()) {
      // input window: rows t0 - pad_left .. + 511 (n_tiles*128 + 127 are used), one 16-byte column of rows
      // per channel chunk; rows outside [0, T) are zero filled by the TMA unit = the conv's zero padding
      mbar_expect_tx(x_full, PC_CHUNKS * PCF_ROWS * 16);
      for (int c = 0; c < PC_CHUNKS; ++c)
        for (int r0 = 0; r0 < PCF_ROWS; r0 += 128)
          tma_load_3d(sX + c * PCF_CS + r0 * 16, &tmX, x_full, g * PC_CG + c * 8, t0 - p.pad_left + r0, b);
      const uint8_t* wg = reinterpret_cast<const uint8_t*>(p.w) + static_cast<size_t>(g) * PC_TAPS * PC_TAP_BYTES;
      int stage = 0;
      uint32_t phase = 0;
      for (int k0 = 0; k0 < PC_TAPS; k0 += PCF_TPS) {
        mbar_wait(&w_empty[stage], phase ^ 1);
        mbar_expect_tx(&w_full[stage], PCF_STAGE_BYTES);
        bulk_load_1d(sW + stage * PCF_STAGE_BYTES, wg + static_cast<size_t>(k0) * PC_TAP_BYTES, PCF_STAGE_BYTES, &w_full[stage]);
        if (++stage == PCF_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, 2 * PC_CG, false, false);
      const uint32_t xa = smem_u32(sX);
      mbar_wait(x_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int k0 = 0; k0 < PC_TAPS; k0 += PCF_TPS) {
        mbar_wait(&w_full[stage], phase);
        tc_fence_after();
        const uint32_t wb = smem_u32(sW + stage * PCF_STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < PCF_TPS / 2; ++kk) {
          const int k = k0 + 2 * kk;  // even tap of the pair
          for (int i = 0; i < n_tiles; ++i) {
#pragma unroll
            for (int ks = 0; ks < PC_CG / 16; ++ks) {
              // A: window rows (i*127 + k) .. +127 of chunks 2ks, 2ks+1;  B: the pair's 96 weight rows of the same chunks
              const uint64_t adesc = make_sdesc_ns(xa + (2 * ks) * PCF_CS + (i * PCF_TSTRIDE + k) * 16, PCF_CS, 128);
              const uint64_t bdesc = make_sdesc_ns(wb + kk * PC_PAIR_BYTES + (2 * ks) * (2 * PC_CG * 16), 2 * PC_CG * 16, 128);
              umma_bf16(tmem_base + i * PCF_TMEM_TILE, adesc, bdesc, idesc, (k > 0 || ks > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(&w_empty[stage]);
        if (++stage == PCF_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------- epilogue: thread = one time row
    const int r = warp * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int i = 0; i < n_tiles; ++i) {
      // D1 (odd taps) belongs to the output row one lane DOWN: fetch this lane's D1 row, hand lane 0's copy to the previous
      // warp through shared memory (its lane 31 needs it), then shift by one lane
      uint32_t d1[48];
      {
        uint32_t lo[32], hi[16];
        tmem_ld32(tmem_base + lane_off + i * PCF_TMEM_TILE + PC_CG, lo);
        tmem_ld16(tmem_base + lane_off + i * PCF_TMEM_TILE + PC_CG + 32, hi);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) d1[c] = lo[c];
#pragma unroll
        for (int c = 0; c < 16; ++c) d1[32 + c] = hi[c];
      }
      float* xch = sXch + (i * 4 + warp) * PC_CG;
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < PC_CG; c += 4)
          *reinterpret_cast<uint4*>(xch + c) = make_uint4(d1[c], d1[c + 1], d1[c + 2], d1[c + 3]);
      }
      bar_sync(1, 128);  // the four epilogue warps (tile i's slots are written once: no second barrier)
#pragma unroll
      for (int c = 0; c < PC_CG; ++c) {
        const uint32_t up = __shfl_down_sync(0xffffffffu, d1[c], 1);
        d1[c] = (lane == 31 && warp < 3) ? __float_as_uint(xch[PC_CG + c]) : up;  // (row 127 has no successor: not stored)
      }
      const int t = t0 + i * PCF_TSTRIDE + r;
      const bool valid = r < PCF_TSTRIDE && t < p.T;  // (no early exit: the TMEM loads below are warp-collective)
      const long long off = (static_cast<long long>(b) * p.T + (valid ? t : 0)) * p.C + g * PC_CG;
#pragma unroll
      for (int c = 0; c < PC_CHUNKS; ++c) {
        uint32_t d0[8];
        tmem_ld8(tmem_base + lane_off + i * PCF_TMEM_TILE + c * 8, d0);
        tmem_ld_wait();
        if (!valid) continue;
        float v[8], res[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(d0[j]) + __uint_as_float(d1[c * 8 + j]);
        bf16x8_to_f32(ldg128(p.res + off + c * 8), res);
        if (p.mode == 0) {
          if (p.bias != nullptr) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + g * PC_CG + c * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + g * PC_CG + c * 8 + 4));
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          }
          // GELU on the half-precision conv output, like nn.GELU under autocast
          const uint4 zq = f32_to_bf16x8(v);
          if (p.z != nullptr) stg128(p.z + off + c * 8, zq);
          bf16x8_to_f32(zq, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += res[j];
        stg128(p.y + off + c * 8, f32_to_bf16x8(v));
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
constexpr int PCW_TT = 256;                         // time rows per pipeline stage
constexpr int PCW_XROWS = PCW_TT + 128;             // 383 used
constexpr int PCW_X_BYTES = PCW_XROWS * 16;         // one channel chunk of x
constexpr int PCW_Y_CS = PCW_TT * 16;               // chunk stride of the dz tile
constexpr int PCW_STAGE_BYTES = PCW_X_BYTES + PC_CHUNKS * PCW_Y_CS;   // 6144 + 24576
constexpr int PCW_STAGES = 4;
constexpr int PCW_SMEM = PCW_STAGES * PCW_STAGE_BYTES + 128 + 128;
constexpr int PCW_THREADS = 192;
constexpr int PCW_JB = PC_TAPS / 16;                // 8 tap blocks of 16 taps -> 8 accumulators [128 x 48]

struct PosConvWgradParams {
  float* dw;   // [C][48][128] fp32, accumulated with red.add
  int B, T, C, splits;
};

__global__ void __launch_bounds__(PCW_THREADS, 1)
posconv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX128,
                     const __grid_constant__ CUtensorMap tmDz, const PosConvWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PCW_STAGES * PCW_STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + PCW_STAGES;
  uint64_t* acc_full = empty + PCW_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x % PC_CHUNKS, g = blockIdx.x / PC_CHUNKS, split = blockIdx.y;
  const int tiles_per_b = (p.T + PCW_TT - 1) / PCW_TT;
  const int n_items = p.B * tiles_per_b;
  const int per = (n_items + p.splits - 1) / p.splits;
  const int it0 = split * per, it1 = min(n_items, it0 + per);

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmX128);
    tma_prefetch_desc(&tmDz);
    for (int s = 0; s < PCW_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = it0; it < it1; ++it) {
        const int b = it / tiles_per_b, t0 = (it % tiles_per_b) * PCW_TT;
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sx = smem + stage * PCW_STAGE_BYTES;
        uint8_t* sy = sx + PCW_X_BYTES;
        mbar_expect_tx(&full[stage], PCW_STAGE_BYTES);
        tma_load_3d(sx, &tmX, &full[stage], g * PC_CG + c * 8, t0 - 64, b);
        tma_load_3d(sx + 256 * 16, &tmX128, &full[stage], g * PC_CG + c * 8, t0 - 64 + 256, b);
        for (int cc = 0; cc < PC_CHUNKS; ++cc)
          tma_load_3d(sy + cc * PCW_Y_CS, &tmDz, &full[stage], g * PC_CG + cc * 8, t0, b);
        if (++stage == PCW_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, PC_CG, true, true);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = it0; it < it1; ++it) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sx = smem_u32(smem + stage * PCW_STAGE_BYTES);
        const uint32_t sy = sx + PCW_X_BYTES;
        for (int ts = 0; ts < PCW_TT / 16; ++ts) {
          // B: dz rows ts*16..+15 (K), 48 out channels (N): chunk stride along N, 8-row groups along K
          // (un-swizzled MN-major operands: the LBO field is the 8-row group stride along K, the SBO field the
          //  16-byte chunk stride along M / N -- verified on hardware against torch's conv weight gradient)
          const uint64_t bdesc = make_sdesc_ns(sy + ts * 256, 128, PCW_Y_CS);
#pragma unroll
          for (int jb = 0; jb < PCW_JB; ++jb) {
            // A: M = (tap j of this block, channel) -> 16 B per tap step, K = time rows
            const uint32_t a0 = sx + (ts * 16 + jb * 16) * 16;
            const uint64_t adesc = make_sdesc_ns(a0, 128, 16);
            umma_bf16(tmem_base + jb * 64, adesc, bdesc, idesc, (it > it0 || ts > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty[stage]);
        if (++stage == PCW_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // epilogue: row m = (tap j = m / 8, channel e = m % 8) of accumulator jb -> dw[g*48 + n][c*8 + e][16 jb + j]
    const int m = warp * 32 + lane;
    const int j = m >> 3, e = m & 7;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (it1 > it0) {
      for (int jb = 0; jb < PCW_JB; ++jb) {
        uint32_t ra[32], rb[16];
        tmem_ld32(tmem_base + lane_off + jb * 64, ra);
        tmem_ld16(tmem_base + lane_off + jb * 64 + 32, rb);
        tmem_ld_wait();
        float* base = p.dw + (static_cast<size_t>(g) * PC_CG * PC_CG + c * 8 + e) * PC_TAPS + jb * 16 + j;
#pragma unroll
        for (int n = 0; n < PC_CG; ++n) {
          const float v = __uint_as_float(n < 32 ? ra[n] : rb[n - 32]);
          atomicAdd(base + static_cast<size_t>(n) * PC_CG * PC_TAPS, v);
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// weight norm (module.py:187: nn.utils.weight_norm(conv, dim=2)): w[:, :, k] = g[k] * v[:, :, k] / ||v[:, :, k]||
// ------------------------------------------------------------------------------------------------
__global__ void pc_normsq_kernel(const float* __restrict__ v, float* __restrict__ normsq, int rows /* C * 48 */) {
  // (one dependent 4-byte load per trip made these three helpers pure load latency -- 60 us each for 19 MB; eight
  //  independent rows per trip now)
  const int k = threadIdx.x;  // 128 taps
  float s = 0.f;
  int r = blockIdx.x;
  for (; r + 7 * static_cast<int>(gridDim.x) < rows; r += 8 * gridDim.x) {
    float x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = __ldg(v + static_cast<size_t>(r + u * gridDim.x) * PC_TAPS + k);
#pragma unroll
    for (int u = 0; u < 8; ++u) s = fmaf(x[u], x[u], s);
  }
  for (; r < rows; r += gridDim.x) {
    const float x = v[static_cast<size_t>(r) * PC_TAPS + k];
    s += x * x;
  }
  atomicAdd(normsq + k, s);
}
// w_fwd[g][k/2][ci/8][(k&1)*48+co][ci%8] = s_k v[g*48+co][ci][k];   w_bwd[g][k'/2][co/8][(k'&1)*48+ci][co%8] = same value, k = 127 - k'
__global__ void pc_weight_prep_kernel(const float* __restrict__ v, const float* __restrict__ gain,
                                      const float* __restrict__ normsq, __nv_bfloat16* __restrict__ w_fwd,
                                      __nv_bfloat16* __restrict__ w_bwd, float* __restrict__ norm_out, int groups) {
  __shared__ float tile[PC_CG][PC_TAPS + 1];  // [ci][k] of one output channel
  const int co_g = blockIdx.x;                // global output channel
  const int g = co_g / PC_CG, co = co_g % PC_CG;
  for (int i = threadIdx.x; i < PC_CG * PC_TAPS; i += blockDim.x) {
    const int ci = i / PC_TAPS, k = i % PC_TAPS;
    const float scale = gain[k] * rsqrtf(normsq[k]);
    tile[ci][k] = v[static_cast<size_t>(co_g) * PC_CG * PC_TAPS + i] * scale;
  }
  if (blockIdx.x == 0 && norm_out != nullptr)
    for (int k = threadIdx.x; k < PC_TAPS; k += blockDim.x) norm_out[k] = sqrtf(normsq[k]);
  __syncthreads();
  for (int i = threadIdx.x; i < PC_CG * PC_TAPS; i += blockDim.x) {
    const int k = i / PC_CG, ci = i % PC_CG;
    const __nv_bfloat16 w = __float2bfloat16(tile[ci][k]);
    // tap-pair layout of posconv_kernel: [g][k / 2][chunk][(k & 1) * 48 + row][8]
    w_fwd[((((static_cast<size_t>(g) * (PC_TAPS / 2) + k / 2) * PC_CHUNKS + ci / 8) * 2 + (k & 1)) * PC_CG + co) * 8 + (ci & 7)] = w;
    if (w_bwd != nullptr) {
      const int kb = PC_TAPS - 1 - k;
      w_bwd[((((static_cast<size_t>(g) * (PC_TAPS / 2) + kb / 2) * PC_CHUNKS + co / 8) * 2 + (kb & 1)) * PC_CG + ci) * 8 + (co & 7)] = w;
    }
  }
}
// dot[k] = sum_{co,ci} dw * v
__global__ void pc_wn_dot_kernel(const float* __restrict__ dw, const float* __restrict__ v, float* __restrict__ dot, int rows) {
  const int k = threadIdx.x;
  float s = 0.f;
  int r = blockIdx.x;
  for (; r + 7 * static_cast<int>(gridDim.x) < rows; r += 8 * gridDim.x) {
    float a[8], b[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const size_t i = static_cast<size_t>(r + u * gridDim.x) * PC_TAPS + k;
      a[u] = __ldg(dw + i);
      b[u] = __ldg(v + i);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) s = fmaf(a[u], b[u], s);
  }
  for (; r < rows; r += gridDim.x) {
    const size_t i = static_cast<size_t>(r) * PC_TAPS + k;
    s += dw[i] * v[i];
  }
  atomicAdd(dot + k, s);
}
// dv += g/n (dw - v dot/n^2);  dg += dot/n
__global__ void pc_wn_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v, const float* __restrict__ gain,
                                 const float* __restrict__ norm, const float* __restrict__ dot, float* __restrict__ dv,
                                 float* __restrict__ dg, int rows) {
  const int k = threadIdx.x;
  const float n = norm[k], gk = gain[k], d = dot[k];
  const float a = gk / n, bcoef = d / (n * n);
  int r = blockIdx.x;
  for (; r + 7 * static_cast<int>(gridDim.x) < rows; r += 8 * gridDim.x) {
    float x[8], y[8], z[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const size_t i = static_cast<size_t>(r + u * gridDim.x) * PC_TAPS + k;
      x[u] = __ldg(dw + i);
      y[u] = __ldg(v + i);
      z[u] = dv[i];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) dv[static_cast<size_t>(r + u * gridDim.x) * PC_TAPS + k] = z[u] + a * (x[u] - y[u] * bcoef);
  }
  for (; r < rows; r += gridDim.x) {
    const size_t i = static_cast<size_t>(r) * PC_TAPS + k;
    dv[i] += a * (dw[i] - v[i] * bcoef);
  }
  if (blockIdx.x == 0) dg[k] += d / n;
}
// dz = dy * gelu'(z)
__global__ void gelu_bwd_mul_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                                    __nv_bfloat16* __restrict__ dz, long long groups) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < groups;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float a[8], b[8];
    bf16x8_to_f32(ldg128(dy + i * 8), a);
    bf16x8_to_f32(ldg128(z + i * 8), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= gelu_erf_grad(b[j]);
    stg128(dz + i * 8, f32_to_bf16x8(a));
  }
}

// x viewed as [B][T][C] bf16, boxes of 8 channels x `box_rows` time rows, no swizzle, zero fill outside [0, T)
static int make_tmap_rows(CUtensorMap* out, const void* base, int B, int T, int C, int box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  MH_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  MH_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && C % 8 == 0, "posconv: tensor must be 16-byte aligned, C %% 8 == 0");
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(T) * C * 2};
  cuuint32_t box[3] = {8, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(posconv rows) failed with %d", static_cast<int>(r));
  return 0;
}
}  // namespace mh

using namespace mh;
#define ST reinterpret_cast<cudaStream_t>(stream)

static int check_shape(int C, int groups, int ktaps) {
  MH_CHECK(groups > 0 && C % groups == 0 && C / groups == PC_CG && ktaps == PC_TAPS,
           "posconv kernels are built for 48 channels per group and 128 taps (got C=%d groups=%d taps=%d)", C, groups, ktaps);
  return 0;
}

extern "C" int mh_posconv_weight_prep(const float* v, const float* g, void* w_fwd, void* w_bwd, float* normsq_ws,
                                      float* norm, int C, int groups, int ktaps, void* stream) {
  if (check_shape(C, groups, ktaps)) return 1;
  MH_CUDA(cudaMemsetAsync(normsq_ws, 0, sizeof(float) * PC_TAPS, ST));
  pc_normsq_kernel<<<296, PC_TAPS, 0, ST>>>(v, normsq_ws, C * PC_CG);
  MH_LAUNCH_CHECK();
  pc_weight_prep_kernel<<<C, 256, 0, ST>>>(v, g, normsq_ws, reinterpret_cast<__nv_bfloat16*>(w_fwd),
                                           reinterpret_cast<__nv_bfloat16*>(w_bwd), norm, groups);
  MH_LAUNCH_CHECK();
  g_launches += 2;
  return 0;
}

static int posconv_launch(const void* x, const void* w, const float* bias, const void* res, void* z, void* y, int B, int T,
                          int C, int pad_left, int mode, cudaStream_t st) {
  CUtensorMap tm;
  int rc = make_tmap_rows(&tm, x, B, T, C, 128);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    MH_CUDA(cudaFuncSetAttribute(posconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PCF_SMEM));
    configured = true;
  }
  PosConvParams p;
  p.w = reinterpret_cast<const __nv_bfloat16*>(w);
  p.bias = bias;
  p.res = reinterpret_cast<const __nv_bfloat16*>(res);
  p.z = reinterpret_cast<__nv_bfloat16*>(z);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.B = B; p.T = T; p.C = C; p.pad_left = pad_left; p.mode = mode;
  const int slabs = (T + PCF_NT * PCF_TSTRIDE - 1) / (PCF_NT * PCF_TSTRIDE);
  posconv_kernel<<<dim3(slabs, C / PC_CG, B), PCF_THREADS, PCF_SMEM, st>>>(tm, p);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

extern "C" int mh_posconv_fwd(const void* x, const void* w_fwd, const float* bias, void* z, void* y, int B, int T, int C,
                              int groups, int ktaps, void* stream) {
  if (check_shape(C, groups, ktaps)) return 1;
  return posconv_launch(x, w_fwd, bias, x, z, y, B, T, C, ktaps / 2, 0, ST);
}
extern "C" int mh_posconv_dgrad(const void* dz, const void* w_bwd, const void* dy, void* dx, int B, int T, int C, int groups,
                                int ktaps, void* stream) {
  if (check_shape(C, groups, ktaps)) return 1;
  return posconv_launch(dz, w_bwd, nullptr, dy, nullptr, dx, B, T, C, ktaps - 1 - ktaps / 2, 1, ST);
}
extern "C" int mh_gelu_bwd_mul(const void* dy, const void* z, void* dz, long long n, void* stream) {
  MH_CHECK(n % 8 == 0, "gelu_bwd_mul: element count must be a multiple of 8");
  long long grid = (n / 8 + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (grid > cap) grid = cap;
  gelu_bwd_mul_kernel<<<static_cast<int>(grid), 256, 0, ST>>>(reinterpret_cast<const __nv_bfloat16*>(dy),
                                                             reinterpret_cast<const __nv_bfloat16*>(z),
                                                             reinterpret_cast<__nv_bfloat16*>(dz), n / 8);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}
extern "C" int mh_posconv_wgrad(const void* dz, const void* x, float* dw, int B, int T, int C, int groups, int ktaps,
                                void* stream) {
  if (check_shape(C, groups, ktaps)) return 1;
  CUtensorMap tx, tx128, tz;
  int rc = make_tmap_rows(&tx, x, B, T, C, 256);
  if (rc) return rc;
  rc = make_tmap_rows(&tx128, x, B, T, C, 128);
  if (rc) return rc;
  rc = make_tmap_rows(&tz, dz, B, T, C, 256);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    MH_CUDA(cudaFuncSetAttribute(posconv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PCW_SMEM));
    configured = true;
  }
  PosConvWgradParams p;
  p.dw = dw; p.B = B; p.T = T; p.C = C;
  const int pairs = groups * PC_CHUNKS;
  const int items = B * ((T + PCW_TT - 1) / PCW_TT);
  int splits = (2 * sm_count() + pairs - 1) / pairs;
  if (splits > items) splits = items;
  if (splits < 1) splits = 1;
  p.splits = splits;
  posconv_wgrad_kernel<<<dim3(pairs, splits), PCW_THREADS, PCW_SMEM, ST>>>(tx, tx128, tz, p);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}
extern "C" int mh_posconv_weight_bwd(const float* dw, const float* v, const float* g, const float* norm, float* dot_ws,
                                     float* dv, float* dg, int C, int groups, int ktaps, void* stream) {
  if (check_shape(C, groups, ktaps)) return 1;
  MH_CUDA(cudaMemsetAsync(dot_ws, 0, sizeof(float) * PC_TAPS, ST));
  pc_wn_dot_kernel<<<296, PC_TAPS, 0, ST>>>(dw, v, dot_ws, C * PC_CG);
  MH_LAUNCH_CHECK();
  pc_wn_bwd_kernel<<<296, PC_TAPS, 0, ST>>>(dw, v, g, norm, dot_ws, dv, dg, C * PC_CG);
  MH_LAUNCH_CHECK();
  g_launches += 2;
  return 0;
}
