// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = A[M,K] * B[N,K]^T   bf16 x bf16 -> fp32 (TMEM) -> fused epilogue -> TMA store
//
// Roles (576 threads, one CTA per SM):
//   warps 0-15 : epilogue, 4 column groups x 4 TMEM lane quadrants.  Warp w owns accumulator rows
//                32*(w%4)..+31 (one row per thread) and the columns of group w/4 (BN/4 wide).  Values are
//                pulled with tcgen05.ld, bias / GELU / dropout / residual / prune-mask are applied in fp32
//                registers, and the result is written into the group's 128-row staging tile in shared
//                memory (hardware swizzle pattern -> conflict-free 16-byte stores).  One elected thread per
//                group then issues a TMA tile store (or fp32 reduce-add for weight gradients), so HBM sees
//                full 128-byte lines and ragged M / N edges are clipped by the hardware.  Residual /
//                pre-activation inputs arrive the same way: TMA load into the staging tile, modified in place.
//   warp 16    : TMA producer (one elected lane): 128B-swizzled tiles of A and B into a 3- or 5-stage
//                shared-memory ring, mbarrier complete_tx signalling.
//   warp 17    : TMEM allocator + MMA issuer (one elected lane): tcgen05.mma 128 x BN x 16, accumulators
//                double buffered in TMEM so the epilogue of tile i overlaps the main loop of tile i+1;
//                tcgen05.commit releases smem stages / publishes tiles.
// Operands may be K-major (row = m or n, 64 contiguous k) or MN-major (row = k, 64 contiguous m or n); the
// latter is what wgrad (dW = dY^T X) needs and avoids any transpose kernel.  MH_EPI_F32 accumulates with
// cp.reduce.async.bulk (.add.f32), which makes split-K and gradient accumulation the same code path.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "mh_b200.h"
#define MH_PDL_FAMILY 1
#include "mh_common.cuh"
#include "mh_ptx.cuh"

#ifndef MH_EPI_KO
#define MH_EPI_KO 0  // timing experiments only (wrong results): 1 no TMA-store read waits, 2 no GELU / dGELU / dropout math
#endif

namespace mh {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int EPI_WARPS = 16;
constexpr int EPI_GROUPS = 4;
constexpr int PRODUCER_WARP = 16;
constexpr int MMA_WARP = 17;
constexpr int GEMM_THREADS = 576;
constexpr int STG_CHUNK = 128 * 128;  // staging bytes per column group: 128 rows x (64 or 128) B

template <int BN>
struct TileCfg {
  static constexpr int kStages = BN == 256 ? 3 : 5;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + EPI_GROUPS * STG_CHUNK + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = 2 * BN;
};

struct GemmDev {
  int M, N, K;
  long long ldd;
  const float* bias;
  int has_aux_out;
  const uint8_t* mask;
  DropCfg drop;
  int split_k;
  float* delta;   // MH_EPI_DELTA: f32 [M / delta_T, N / 64, delta_T]
  int delta_T;
};

// byte offset of 16-byte chunk `ch` of row `r` in a staging tile whose rows are ROWB bytes, laid out the
// way TMA expects for CU_TENSOR_MAP_SWIZZLE_128B (ROWB = 128) / SWIZZLE_64B (ROWB = 64)
template <int ROWB>
__device__ __forceinline__ uint32_t stg_off(int r, int ch) {
  if (ROWB == 128) return r * 128 + ((ch ^ (r & 7)) << 4);
  return r * 64 + ((ch ^ ((r >> 1) & 3)) << 4);
}

// ---------------------------------------------------------------------------------------------
// Epilogue of one output tile (128 accumulator rows of this CTA x BN columns), shared by the 1-CTA and the
// CTA-pair kernels.  Executed by the 16 epilogue warps; see the header comment for the data flow.
// ---------------------------------------------------------------------------------------------
struct EpiThread {
  int quad, grp, lane, r_tile, bar_id;
  bool leader;
  uint8_t* stg;
  uint32_t lane_off;
};

__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmDev& p, const CUtensorMap* tmD, const CUtensorMap* tmAuxIn,
                                              const CUtensorMap* tmAuxOut, const EpiThread& e, uint32_t tmem_acc, int m0,
                                              int n0, bool has_k, uint64_t* tfull_bar, uint32_t acc_phase, uint64_t* aux_bar,
                                              uint32_t& aux_phase, uint32_t tempty_cluster_addr) {
  constexpr bool OUT_F32 = EPI == MH_EPI_F32;
  constexpr bool HAS_AUX_IN = EPI == MH_EPI_RES || EPI == MH_EPI_DGELU || EPI == MH_EPI_ADD || EPI == MH_EPI_DELTA;
  constexpr int CG = BN / EPI_GROUPS;                    // columns per epilogue group
  constexpr int SUBC = OUT_F32 ? 32 : CG;                // columns staged per pass (a staging row is <= 128 bytes)
  constexpr int ROWB = SUBC * (OUT_F32 ? 4 : 2);         // staging row bytes
  static_assert(ROWB == 64 || ROWB == 128, "staging rows are 64 or 128 bytes");
  const int gcol0 = n0 + e.grp * CG;
  const bool active = gcol0 < p.N && has_k;  // uniform over the group
  const long long row = m0 + e.r_tile;
  uint8_t* stg = e.stg;

  // (1) the staging tile is free once the previous TMA store has read it; residual / pre-activation
  //     tiles are fetched into it right away so the load overlaps the wait for the accumulator
  if (e.leader) {
    if (!(MH_EPI_KO & 1)) bulk_wait_read0();
    if (HAS_AUX_IN && active) {
      mbar_expect_tx(aux_bar, 128 * ROWB);
      tma_load_2d(stg, tmAuxIn, aux_bar, gcol0, m0);
    }
  }
  if (HAS_AUX_IN) {
    if (active) {
      mbar_wait(aux_bar, aux_phase);
      aux_phase ^= 1;
    }
  } else {
    bar_sync(e.bar_id, 128);
  }

  // (2) accumulator -> registers -> staging
  mbar_wait(tfull_bar, acc_phase);
  tc_fence_after();
  if (OUT_F32) {
    // fp32 reduce-add (weight gradients): 32 columns per pass through the 16 KB staging tile
#pragma unroll 1
    for (int s = 0; s < CG / 32; ++s) {
      const int col_in_tile = e.grp * CG + s * 32;
      const bool sub_active = active && (n0 + col_in_tile) < p.N;
      if (s > 0) {
        if (e.leader) bulk_wait_read0();
        bar_sync(e.bar_id, 128);
      }
      if (sub_active) {
        uint32_t r[32];
        tmem_ld32(tmem_acc + e.lane_off + col_in_tile, r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = n0 + col_in_tile + q * 8;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[q * 8 + j]);
          if (p.mask != nullptr && col < p.N && row < p.M) {
            const uint2 mk = *reinterpret_cast<const uint2*>(p.mask + row * p.ldd + col);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (((mk.x >> (8 * j)) & 0xFF) == 0) v[j] = 0.f;
              if (((mk.y >> (8 * j)) & 0xFF) == 0) v[4 + j] = 0.f;
            }
          }
          *reinterpret_cast<float4*>(stg + stg_off<ROWB>(e.r_tile, 2 * q)) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(stg + stg_off<ROWB>(e.r_tile, 2 * q + 1)) = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
      if (s == CG / 32 - 1) {  // last TMEM read of this tile: hand the accumulator back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (e.lane == 0) mbar_arrive_cluster(tempty_cluster_addr);
      }
      if (sub_active) {
        fence_proxy_async_smem();
        bar_sync(e.bar_id, 128);
        if (e.leader) {
          tma_reduce_add_2d(tmD, stg, n0 + col_in_tile, m0);
          bulk_commit();
        }
      }
    }
    return;
  }

  // GELU with two outputs: activated values wait in registers while `pre` is stored.  The 32-column
  // sub-chunk loop stays rolled (the unrolled 64-column body was ~30 KB of SASS per kernel and fetch-bound
  // like the attention loops), so the held values live in two explicitly named register sets.
  uint4 held0[4], held1[4];
  // dropout stream position of this thread's row (element (row, col) lives in word row * N/32 + col/32, bit col % 32)
  const uint64_t drop_base = static_cast<uint64_t>(row) * static_cast<uint64_t>(p.N >> 5) + static_cast<uint64_t>(n0 >> 5);
  const DropState dstate(p.drop);  // (device step counter: one L1-resident load per tile)
  const uint32_t stg_s = smem_u32(stg);
  const float drop_s = (EPI == MH_EPI_GELU && p.drop.thresh != 0) ? p.drop.scale : 1.f;
  const float drop_lg = (EPI == MH_EPI_GELU && p.drop.thresh != 0) ? log2f(p.drop.scale) : 0.f;
  float dsum = 0.f;  // MH_EPI_DELTA: this row's dot product over the group's 64 columns (= one attention head)
  if (active) {
    auto sub_chunk = [&](int s) {
      const int col_in_tile = e.grp * CG + s * 32;
      // the 32 keep decisions of this thread's 32 columns: generated while the accumulator load is in flight
      constexpr bool HAS_DROP = EPI == MH_EPI_GELU || EPI == MH_EPI_RES || EPI == MH_EPI_DGELU;
      uint32_t r[32];
      tmem_ld32(tmem_acc + e.lane_off + col_in_tile, r);
      uint32_t kw = 0xffffffffu;
      if (HAS_DROP && p.drop.thresh != 0 && !(MH_EPI_KO & 2)) kw = dstate.keep32(p.drop, drop_base + (col_in_tile >> 5));
      tmem_ld_wait();
      uint4 o4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = n0 + col_in_tile + q * 8;
        const bool col_ok = col < p.N;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[q * 8 + j]);
        const int ch = s * 4 + q;  // 16-byte chunk of the staging row
        uint8_t* slot = stg + stg_off<ROWB>(e.r_tile, ch);
        const uint32_t slot_s = stg_s + stg_off<ROWB>(e.r_tile, ch);  // (shared-window address: STS / LDS, no generic ST.E)
        if (EPI == MH_EPI_BF16 || EPI == MH_EPI_GELU || EPI == MH_EPI_RES) {
          if (p.bias != nullptr && col_ok) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
            unpack2f(fadd2(pack2f(v[0], v[1]), pack2f(b0.x, b0.y)), v[0], v[1]);
            unpack2f(fadd2(pack2f(v[2], v[3]), pack2f(b0.z, b0.w)), v[2], v[3]);
            unpack2f(fadd2(pack2f(v[4], v[5]), pack2f(b1.x, b1.y)), v[4], v[5]);
            unpack2f(fadd2(pack2f(v[6], v[7]), pack2f(b1.z, b1.w)), v[6], v[7]);
          }
        }
        if (EPI == MH_EPI_GELU) {
          // the reference evaluates GELU in fp32 on the half-precision fc1 output
          // (fairseq_code/gelu.py:35 under autocast): round first, then activate.
          const uint4 pre = f32_to_bf16x8(v);
          if (p.has_aux_out) sts128(slot_s, pre);
          bf16x8_to_f32(pre, v);
          if (!(MH_EPI_KO & 2)) gelu_erf8(v, drop_s, drop_lg);  // (the dropout keep-scale rides along in the exponent: no multiply later)
        }
        if (EPI == MH_EPI_DGELU) {
          float pre[8];
          bf16x8_to_f32(lds128_plain(slot_s), pre);
          if (!(MH_EPI_KO & 2)) gelu_erf_grad_mul8(v, pre);
        }
        if (HAS_DROP) {
          if (p.drop.thresh != 0) {
            if (EPI == MH_EPI_GELU) {  // already scaled
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = ((kw >> (8 * q)) & (1u << j)) ? v[j] : 0.f;
            } else {
              DropState::apply8(p.drop, kw >> (8 * q), v);
            }
          }
        }
        if (EPI == MH_EPI_RES || EPI == MH_EPI_ADD) {
          float a[8];
          bf16x8_to_f32(lds128_plain(slot_s), a);
#pragma unroll
          for (int j = 0; j < 4; ++j) unpack2f(fadd2(pack2f(v[2 * j], v[2 * j + 1]), pack2f(a[2 * j], a[2 * j + 1])), v[2 * j], v[2 * j + 1]);
        }
        o4[q] = f32_to_bf16x8(v);
        if (EPI == MH_EPI_DELTA) {
          // delta = rowsum(dO * O) per head (the softmax-backward correction term of the attention backward): dO is this
          // GEMM's (rounded) output, the O tile was prefetched into the staging tile like a residual
          float g[8], o[8];
          bf16x8_to_f32(o4[q], g);
          bf16x8_to_f32(lds128_plain(slot_s), o);
#pragma unroll
          for (int j = 0; j < 8; ++j) dsum = fmaf(g[j], o[j], dsum);
        }
        if (!(EPI == MH_EPI_GELU && p.has_aux_out)) sts128(slot_s, o4[q]);
      }
      if (EPI == MH_EPI_GELU && p.has_aux_out) {
        if (s == 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) held0[q] = o4[q];
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) held1[q] = o4[q];
        }
      }
    };
    if (EPI == MH_EPI_GELU || EPI == MH_EPI_DGELU) {
      // heavy epilogues: rolled (measured 146 -> 132 us for fc1 + GELU + dropout)
#pragma unroll 1
      for (int s = 0; s < CG / 32; ++s) sub_chunk(s);
    } else {
      // light epilogues: unrolled (the loop overhead costs more than the extra code)
#pragma unroll
      for (int s = 0; s < CG / 32; ++s) sub_chunk(s);
    }
  }
  if (EPI == MH_EPI_DELTA && active && row < p.M) {
    static_assert(EPI != MH_EPI_DELTA || CG == 64, "the delta epilogue needs 64-column groups (one head each): BN = 256");
    const long long b = row / p.delta_T, t = row - b * p.delta_T;
    p.delta[(b * (p.N >> 6) + (gcol0 >> 6)) * p.delta_T + t] = dsum;
  }
  // (3) TMEM buffer back to the MMA warp
  tc_fence_before();
  __syncwarp();
  if (e.lane == 0) mbar_arrive_cluster(tempty_cluster_addr);

  // (4) staging -> global
  if (active) {
    fence_proxy_async_smem();
    bar_sync(e.bar_id, 128);
    if (EPI == MH_EPI_GELU && p.has_aux_out) {
      if (e.leader) {
        tma_store_2d(tmAuxOut, stg, gcol0, m0);
        bulk_commit();
        if (!(MH_EPI_KO & 1)) bulk_wait_read0();
      }
      bar_sync(e.bar_id, 128);
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(stg + stg_off<ROWB>(e.r_tile, q)) = held0[q];
      if (CG > 32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(stg + stg_off<ROWB>(e.r_tile, 4 + q)) = held1[q];
      }
      fence_proxy_async_smem();
      bar_sync(e.bar_id, 128);
    }
    if (e.leader) {
      tma_store_2d(tmD, stg, gcol0, m0);
      bulk_commit();
    }
  }
}

__device__ __forceinline__ EpiThread make_epi_thread(int warp, int lane, uint8_t* staging) {
  EpiThread e;
  e.quad = warp & 3;   // TMEM lane quadrant
  e.grp = warp >> 2;   // column group
  e.lane = lane;
  e.r_tile = e.quad * 32 + lane;
  e.bar_id = 1 + e.grp;
  e.leader = e.quad == 0 && lane == 0;
  e.stg = staging + e.grp * STG_CHUNK;
  e.lane_off = static_cast<uint32_t>(e.quad * 32) << 16;
  return e;
}

template <int BN, int EPI, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmAuxIn,
            const __grid_constant__ CUtensorMap tmAuxOut, const GemmDev p) {
  pdl_launch_dependents();  // (pdl_wait() follows the set-up below: nothing in front of it reads global memory)
  using Cfg = TileCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + EPI_GROUPS * STG_CHUNK);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tfull = bars + 2 * Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* aux_full = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_full + EPI_GROUPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == PRODUCER_WARP && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], EPI_WARPS);
    }
    for (int s = 0; s < EPI_GROUPS; ++s) mbar_init(&aux_full[s], 1);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  pdl_wait();  // barriers initialised, TMEM allocated, descriptors prefetched under the previous kernel's tail
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m = (p.M + BM - 1) / BM;
  const int num_n = (p.N + BN - 1) / BN;
  const int kblocks_total = (p.K + BK - 1) / BK;
  const int splits = p.split_k;
  const int kb_per_split = (kblocks_total + splits - 1) / splits;
  const int num_tiles = num_m * num_n * splits;

  if (warp == PRODUCER_WARP) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int split = tile % splits;
        const int mn = tile / splits;
        const int m0 = (mn / num_n) * BM, n0 = (mn % num_n) * BN;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kblocks_total, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], Cfg::kStageBytes);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (!A_MN) {
            tma_load_2d(sa, &tmA, &full[stage], kb * BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * (BK * 128), &tmA, &full[stage], m0 + c * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, &full[stage], kb * BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * (BK * 128), &tmB, &full[stage], n0 + c * 64, kb * BK);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int split = tile % splits;
        const int mn = tile / splits;
        const int n0 = (mn % num_n) * BN;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kblocks_total, kb0 + kb_per_split);
        // shrink the instruction N on the ragged last column tile (multiples of 16)
        int n_valid = min(BN, p.N - n0);
        n_valid = (n_valid + 15) & ~15;
        const uint32_t idesc = make_idesc_bf16(BM, n_valid, A_MN, B_MN);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_MN ? make_sdesc(sa + k * 2048, BK * 128, 1024) : make_sdesc(sa + k * 32, 0, 1024);
            const uint64_t bdesc = B_MN ? make_sdesc(sb + k * 2048, BK * 128, 1024) : make_sdesc(sb + k * 32, 0, 1024);
            umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    const EpiThread e = make_epi_thread(warp, lane, staging);
    int acc = 0;
    uint32_t acc_phase = 0, aux_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mn = tile / splits;
      const int m0 = (mn / num_n) * BM, n0 = (mn % num_n) * BN;
      const int kb0 = (tile % splits) * kb_per_split;
      const bool has_k = kb0 < kblocks_total;  // empty split (possible when splits does not divide)
      epilogue_tile<BN, EPI>(p, &tmD, &tmAuxIn, &tmAuxOut, e, tmem_base + acc * BN, m0, n0, has_k, &tfull[acc], acc_phase,
                             &aux_full[e.grp], aux_phase, smem_u32(&tempty[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (e.leader) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of a cluster compute one 256 x 256 tile.  Each CTA loads
// its own 128 rows of A and HALF of the B tile (128 of the 256 columns); the leader CTA issues one
// 256 x 256 x 16 MMA per k-step that reads both CTAs' shared memory and writes 128 accumulator rows into
// each CTA's TMEM.  Per SM this halves the B-operand shared-memory traffic (the 1-CTA kernel sits at the
// 128 B/clk shared-memory roof once the epilogue staging traffic is added) and the smaller stage (32 KB)
// buys a 5-deep ring.  Each CTA runs the normal epilogue on its own 128 rows.
//   full[s]   : leader's barrier; both CTAs' TMA loads complete_tx on it (peer bit cleared in the address)
//   empty[s]  : one per CTA; released by a multicast tcgen05.commit
//   tfull[a]  : one per CTA (multicast commit);  tempty[a]: leader's, 2 x 16 warp arrivals (remote for the peer)
// ---------------------------------------------------------------------------------------------
constexpr int G2_BN = 256;
constexpr int G2_STAGES = 5;
constexpr int G2_A_BYTES = BM * BK * 2;          // this CTA's 128 rows of A
constexpr int G2_B_BYTES = (G2_BN / 2) * BK * 2; // this CTA's half of the B tile
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_SMEM = G2_STAGES * G2_STAGE_BYTES + EPI_GROUPS * STG_CHUNK + 1024 + 256;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;      // clears the CTA-rank bit of a shared::cluster address -> even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {  // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
               : "memory");
}

template <int EPI, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmAuxIn,
                 const __grid_constant__ CUtensorMap tmAuxOut, const GemmDev p) {
  pdl_launch_dependents();  // (pdl_wait() follows the set-up below: nothing in front of it reads global memory)
  constexpr int BN = G2_BN;
  extern __shared__ uint8_t smem_raw[];
  // identical offsets in both CTAs (the dynamic smem base is the same for every CTA of a launch)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + G2_STAGES * G2_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + EPI_GROUPS * STG_CHUNK);
  uint64_t* full = bars;
  uint64_t* empty = bars + G2_STAGES;
  uint64_t* tfull = bars + 2 * G2_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* aux_full = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_full + EPI_GROUPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader = rank == 0;

  if (warp == PRODUCER_WARP && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 2 * EPI_WARPS);
    }
    for (int s = 0; s < EPI_GROUPS; ++s) mbar_init(&aux_full[s], 1);
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  pdl_wait();  // barriers initialised, TMEM allocated, descriptors prefetched under the previous kernel's tail
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_m2 = (p.M + 2 * BM - 1) / (2 * BM);
  const int num_n = (p.N + BN - 1) / BN;
  const int kblocks_total = (p.K + BK - 1) / BK;
  const int splits = p.split_k;
  const int kb_per_split = (kblocks_total + splits - 1) / splits;
  const int num_tiles = num_m2 * num_n * splits;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == PRODUCER_WARP) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int split = tile % splits;
        const int mn = tile / splits;
        const int m0 = (mn / num_n) * (2 * BM) + rank * BM;
        const int n0 = (mn % num_n) * BN + rank * (BN / 2);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kblocks_total, kb0 + kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t fb = smem_u32(&full[stage]) & PEER_MASK;  // the leader's barrier
          if (is_leader) mbar_expect_tx(&full[stage], 2 * G2_STAGE_BYTES);
          uint8_t* sa = smem + stage * G2_STAGE_BYTES;
          uint8_t* sb = sa + G2_A_BYTES;
          if (!A_MN) {
            tma_load_2d_pair(sa, &tmA, fb, kb * BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d_pair(sa + c * (BK * 128), &tmA, fb, m0 + c * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d_pair(sb, &tmB, fb, kb * BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 128; ++c) tma_load_2d_pair(sb + c * (BK * 128), &tmB, fb, n0 + c * 64, kb * BK);
          }
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    if (is_leader && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = make_idesc_bf16(2 * BM, BN, A_MN, B_MN);
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int split = tile % splits;
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kblocks_total, kb0 + kb_per_split);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * G2_STAGE_BYTES);
          const uint32_t sb = sa + G2_A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = A_MN ? make_sdesc(sa + k * 2048, BK * 128, 1024) : make_sdesc(sa + k * 32, 0, 1024);
            const uint64_t bdesc = B_MN ? make_sdesc(sb + k * 2048, BK * 128, 1024) : make_sdesc(sb + k * 32, 0, 1024);
            umma_bf16_pair(tmem_d, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty[stage]);
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    const EpiThread e = make_epi_thread(warp, lane, staging);
    int acc = 0;
    uint32_t acc_phase = 0, aux_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int mn = tile / splits;
      const int m0 = (mn / num_n) * (2 * BM) + rank * BM, n0 = (mn % num_n) * BN;
      const int kb0 = (tile % splits) * kb_per_split;
      const bool has_k = kb0 < kblocks_total;
      epilogue_tile<BN, EPI>(p, &tmD, &tmAuxIn, &tmAuxOut, e, tmem_base + acc * BN, m0, n0, has_k, &tfull[acc], acc_phase,
                             &aux_full[e.grp], aux_phase, smem_u32(&tempty[acc]) & PEER_MASK);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (e.leader) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's shared memory / barriers must stay alive until both CTAs are done
  if (warp == MMA_WARP)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// Descriptor cache ("the library owns only immutable caches: TMA descriptors keyed by ptr / shape / stride").
// An eager training step encodes ~750 tensor maps on the host (5 per GEMM launch) for the same few hundred (pointer,
// shape) pairs every step -- activations come back from the caching allocator at the same addresses; a CUtensorMap is a
// pure function of its key, so entries never go stale.  Behind a mutex: backward runs on autograd's worker thread.
struct TmapKey {
  const void* base;
  long long d[5];
  int b[4];
  bool operator==(const TmapKey& o) const { return base == o.base && !memcmp(d, o.d, sizeof(d)) && !memcmp(b, o.b, sizeof(b)); }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
    for (long long v : k.d) h = (h ^ static_cast<uint64_t>(v)) * 0x100000001B3ull;
    for (int v : k.b) h = (h ^ static_cast<uint64_t>(static_cast<uint32_t>(v))) * 0x100000001B3ull;
    return static_cast<size_t>(h ^ (h >> 29));
  }
};
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static bool tmap_lookup(const TmapKey& k, CUtensorMap* out) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  auto it = g_tmap_cache.find(k);
  if (it == g_tmap_cache.end()) return false;
  *out = it->second;
  return true;
}
static void tmap_insert(const TmapKey& k, const CUtensorMap& m) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (g_tmap_cache.size() >= 8192) g_tmap_cache.clear();  // bounded: a long run with changing shapes starts over
  g_tmap_cache.emplace(k, m);
}

// 2-D bf16 row-major array [rows][cols] (leading dim ld elements), box = {box_cols, box_rows},
// 128-byte swizzle, out-of-bounds reads return zeros.
int make_tmap_2d(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld, int box_cols,
                 int box_rows, int elem_bytes = 2, int swizzle_bytes = 128) {
  const TmapKey key{base, {rows, cols, ld, 0, 2}, {box_cols, box_rows, elem_bytes, swizzle_bytes}};
  if (tmap_lookup(key, out)) return 0;
  EncodeTiledFn enc = get_encode_tiled();
  MH_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  MH_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  MH_CHECK((ld * elem_bytes) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes (ld = %lld elements)", ld);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (rows=%lld cols=%lld ld=%lld box=%dx%d)",
           static_cast<int>(r), rows, cols, ld, box_cols, box_rows);
  tmap_insert(key, *out);
  return 0;
}

// 3-D variant used by attention: [d2][d1][d0] with strides in elements.
int make_tmap_3d(CUtensorMap* out, const void* base, long long d0, long long d1, long long d2, long long stride1,
                 long long stride2, int box0, int box1, int elem_bytes) {
  const TmapKey key{base, {d0, d1, d2, stride1, stride2 ^ (3LL << 60)}, {box0, box1, elem_bytes, 128}};
  if (tmap_lookup(key, out)) return 0;
  EncodeTiledFn enc = get_encode_tiled();
  MH_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  MH_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  MH_CHECK((stride1 * elem_bytes) % 16 == 0 && (stride2 * elem_bytes) % 16 == 0, "TMA strides must be multiples of 16 bytes");
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2)};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(stride1) * elem_bytes, static_cast<cuuint64_t>(stride2) * elem_bytes};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box0), static_cast<cuuint32_t>(box1), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MH_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed with %d", static_cast<int>(r));
  tmap_insert(key, *out);
  return 0;
}

extern long long g_launches;

struct GemmMaps {
  CUtensorMap a, b, d, aux_in, aux_out;
};

template <int BN, int EPI, bool A_MN, bool B_MN>
static int launch(const GemmMaps& t, const GemmDev& d, int grid, cudaStream_t st) {
  auto kfn = gemm_kernel<BN, EPI, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    MH_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, TileCfg<BN>::kSmemBytes));
    configured = true;
  }
  MH_CUDA(launch_pdl(kfn, dim3(grid), dim3(GEMM_THREADS), TileCfg<BN>::kSmemBytes, st, t.a, t.b, t.d, t.aux_in, t.aux_out, d));
  ++g_launches;
  return 0;
}

template <int EPI, bool A_MN, bool B_MN>
static int launch_pair(const GemmMaps& t, const GemmDev& d, int grid, cudaStream_t st) {
  auto kfn = gemm_pair_kernel<EPI, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    MH_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM));
    configured = true;
  }
  MH_CUDA(launch_pdl(kfn, dim3(grid), dim3(GEMM_THREADS), G2_SMEM, st, t.a, t.b, t.d, t.aux_in, t.aux_out, d));
  ++g_launches;
  return 0;
}

template <int EPI>
static int dispatch_major_pair(const mh_gemm_args* a, const GemmMaps& t, const GemmDev& d, int grid, cudaStream_t st) {
  if (!a->a_mn && !a->b_mn) return launch_pair<EPI, false, false>(t, d, grid, st);
  if constexpr (EPI == MH_EPI_BF16 || EPI == MH_EPI_DGELU || EPI == MH_EPI_ADD || EPI == MH_EPI_F32 || EPI == MH_EPI_DELTA) {
    if (!a->a_mn && a->b_mn) return launch_pair<EPI, false, true>(t, d, grid, st);
  }
  if constexpr (EPI == MH_EPI_F32 || EPI == MH_EPI_BF16) {
    if (a->a_mn && a->b_mn) return launch_pair<EPI, true, true>(t, d, grid, st);
    if (a->a_mn && !a->b_mn) return launch_pair<EPI, true, false>(t, d, grid, st);
  }
  set_error("operand layout a_mn=%d b_mn=%d is not built for epilogue %d", a->a_mn, a->b_mn, a->epilogue);
  return 1;
}

static int dispatch_epi_pair(const mh_gemm_args* a, const GemmMaps& t, const GemmDev& d, int grid, cudaStream_t st) {
  switch (a->epilogue) {
    case MH_EPI_BF16: return dispatch_major_pair<MH_EPI_BF16>(a, t, d, grid, st);
    case MH_EPI_GELU: return dispatch_major_pair<MH_EPI_GELU>(a, t, d, grid, st);
    case MH_EPI_RES: return dispatch_major_pair<MH_EPI_RES>(a, t, d, grid, st);
    case MH_EPI_F32: return dispatch_major_pair<MH_EPI_F32>(a, t, d, grid, st);
    case MH_EPI_DGELU: return dispatch_major_pair<MH_EPI_DGELU>(a, t, d, grid, st);
    case MH_EPI_ADD: return dispatch_major_pair<MH_EPI_ADD>(a, t, d, grid, st);
    case MH_EPI_DELTA: return dispatch_major_pair<MH_EPI_DELTA>(a, t, d, grid, st);
  }
  set_error("unknown epilogue %d", a->epilogue);
  return 1;
}

template <int BN, int EPI>
static int dispatch_major(const mh_gemm_args* a, const GemmMaps& t, const GemmDev& d, int grid, cudaStream_t st) {
  if constexpr (EPI == MH_EPI_F32 && BN != 128) {
    set_error("fp32 accumulate epilogue is built for block_n = 128 only");
    return 1;
  } else if constexpr (EPI == MH_EPI_DELTA) {
    if constexpr (BN == 256) {
      if (!a->a_mn && a->b_mn) return launch<BN, EPI, false, true>(t, d, grid, st);
    }
    set_error("the delta epilogue is built for block_n = 256, a K-major, b MN-major only");
    return 1;
  } else {
    if (!a->a_mn && !a->b_mn) return launch<BN, EPI, false, false>(t, d, grid, st);
    // dgrad reads the forward weight [N][K] as an MN-major B operand: plain / +add / dGELU outputs
    if constexpr (EPI == MH_EPI_BF16 || EPI == MH_EPI_DGELU || EPI == MH_EPI_ADD || EPI == MH_EPI_F32) {
      if (!a->a_mn && a->b_mn) return launch<BN, EPI, false, true>(t, d, grid, st);
    }
    // wgrad (dW = dY^T X) reads both activations MN-major
    if constexpr (EPI == MH_EPI_F32 || EPI == MH_EPI_BF16) {
      if (a->a_mn && a->b_mn) return launch<BN, EPI, true, true>(t, d, grid, st);
      if (a->a_mn && !a->b_mn) return launch<BN, EPI, true, false>(t, d, grid, st);
    }
    set_error("operand layout a_mn=%d b_mn=%d is not built for epilogue %d", a->a_mn, a->b_mn, a->epilogue);
    return 1;
  }
}

template <int BN>
static int dispatch_epi(const mh_gemm_args* a, const GemmMaps& t, const GemmDev& d, int grid, cudaStream_t st) {
  switch (a->epilogue) {
    case MH_EPI_BF16: return dispatch_major<BN, MH_EPI_BF16>(a, t, d, grid, st);
    case MH_EPI_GELU: return dispatch_major<BN, MH_EPI_GELU>(a, t, d, grid, st);
    case MH_EPI_RES: return dispatch_major<BN, MH_EPI_RES>(a, t, d, grid, st);
    case MH_EPI_F32: return dispatch_major<BN, MH_EPI_F32>(a, t, d, grid, st);
    case MH_EPI_DGELU: return dispatch_major<BN, MH_EPI_DGELU>(a, t, d, grid, st);
    case MH_EPI_ADD: return dispatch_major<BN, MH_EPI_ADD>(a, t, d, grid, st);
    case MH_EPI_DELTA: return dispatch_major<BN, MH_EPI_DELTA>(a, t, d, grid, st);
  }
  set_error("unknown epilogue %d", a->epilogue);
  return 1;
}

}  // namespace mh

extern "C" int mh_gemm(const mh_gemm_args* a, void* stream) {
  using namespace mh;
  MH_CHECK(a != nullptr, "null args");
  MH_CHECK(a->M > 0 && a->N > 0 && a->K > 0, "bad GEMM shape %d x %d x %d", a->M, a->N, a->K);
  MH_CHECK(a->N % 8 == 0, "N must be a multiple of 8 (got %d)", a->N);
  MH_CHECK(a->ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(a->D) & 15) == 0, "D must be 16-byte aligned, ldd %% 8 == 0");
  if (a->epilogue == MH_EPI_RES || a->epilogue == MH_EPI_DGELU || a->epilogue == MH_EPI_ADD || a->epilogue == MH_EPI_DELTA)
    MH_CHECK(a->aux_in != nullptr && a->ld_aux % 8 == 0, "epilogue %d needs aux_in", a->epilogue);
  if (a->epilogue == MH_EPI_DELTA)
    MH_CHECK(a->delta != nullptr && a->delta_T > 0 && a->M % a->delta_T == 0 && a->N % 64 == 0 && !a->a_mn && a->b_mn,
             "delta epilogue: needs delta, delta_T dividing M, N %% 64 == 0, a K-major and b MN-major");
  if (a->epilogue == MH_EPI_GELU && a->aux_out != nullptr) MH_CHECK(a->ld_aux % 8 == 0, "ld_aux %% 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  const int sms = sm_count();
  const bool out_f32 = a->epilogue == MH_EPI_F32;
  const int es = out_f32 ? 4 : 2;
  MH_CHECK((a->ldd * es) % 16 == 0, "D row pitch must be a multiple of 16 bytes");
  const int num_m = (a->M + BM - 1) / BM;
  const int kblocks = (a->K + BK - 1) / BK;
  // tile shape: block_n 0 = auto, 128 / 256 = single-CTA 128 x block_n tiles, -256 = CTA-pair 256 x 256 tiles
  int bn = a->block_n;
  static const bool pair_allowed = [] { const char* e = getenv("MH_GEMM_PAIR"); return !(e && e[0] == '0'); }();
  bool pair = false;
  if (bn == -256) {
    pair = true;
  } else if (bn == 0 && pair_allowed && a->M >= 2 * BM && a->N >= 128) {
    // big GEMMs: the pair kernel whenever its 256 x 256 tiles (x split-K for wgrad) can occupy the SM pairs
    const int pt = ((a->M + 255) / 256) * ((a->N + 255) / 256);
    pair = out_f32 ? true : pt >= sms / 2;
  }
  if (!pair) {
    if (out_f32) {
      MH_CHECK(bn == 0 || bn == 128, "fp32 accumulate epilogue: block_n must be 0, 128 or -256");
      bn = 128;
    }
    if (bn == 0) bn = (a->N >= 256 && num_m * ((a->N + 255) / 256) >= sms) ? 256 : 128;
    if (a->epilogue == MH_EPI_DELTA) bn = 256;  // 64-column epilogue groups = one attention head each
    MH_CHECK(bn == 128 || bn == 256, "block_n must be 0, 128, 256 or -256");
  } else {
    bn = G2_BN;
  }
  const int tiles_m = pair ? (a->M + 2 * BM - 1) / (2 * BM) : num_m;
  const int num_n = (a->N + bn - 1) / bn;
  const int units = pair ? sms / 2 : sms;  // CTAs or CTA pairs that can be resident
  int splits = 1;
  if (out_f32) {
    splits = a->split_k;
    if (splits <= 0) {
      // Work units = tiles x splits, `units` of them run at a time: pick the split count that minimises
      // waves x (k-blocks per unit + a fixed per-unit cost: pipeline fill and the part of the reduce-add epilogue the next
      // unit's main loop does not hide, ~6 k-blocks).  "units / tiles" alone left the fused QKV weight gradient (27 tiles on
      // 74 CTA pairs) at 2 splits = 54 units = 73 % of the chip; 8 splits = 216 units = 2.92 waves.
      const int tiles_mn = tiles_m * num_n;
      const int max_s = kblocks / 4 > 0 ? (kblocks / 4 < 32 ? kblocks / 4 : 32) : 1;  // keep >= 4 k-blocks per split
      long best_cost = -1;
      splits = 1;
      for (int sp = 1; sp <= max_s; ++sp) {
        const int per = (kblocks + sp - 1) / sp;
        const int eff = (kblocks + per - 1) / per;  // no empty trailing splits
        const long waves = (static_cast<long>(tiles_mn) * eff + units - 1) / units;
        const long cost = waves * (per + 6);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; splits = eff; }
      }
    }
    if (splits > kblocks) splits = kblocks;
    // avoid empty trailing splits
    const int per = (kblocks + splits - 1) / splits;
    splits = (kblocks + per - 1) / per;
  }

  GemmMaps t;
  int rc;
  const int b_rows = pair ? bn / 2 : bn;  // B rows (K-major) each CTA loads per stage
  if (!a->a_mn) rc = make_tmap_2d(&t.a, a->A, a->M, a->K, a->lda, BK, BM);
  else rc = make_tmap_2d(&t.a, a->A, a->K, a->M, a->lda, 64, BK);
  if (rc) return rc;
  if (!a->b_mn) rc = make_tmap_2d(&t.b, a->B, a->N, a->K, a->ldb, BK, b_rows);
  else rc = make_tmap_2d(&t.b, a->B, a->K, a->N, a->ldb, 64, BK);
  if (rc) return rc;
  // epilogue tiles: one box of 128 rows x (bn / 4) columns per column group (fp32: 32-column passes)
  const int cg = out_f32 ? 32 : bn / EPI_GROUPS;
  const int rowb = cg * es;
  rc = make_tmap_2d(&t.d, a->D, a->M, a->N, a->ldd, cg, BM, es, rowb);
  if (rc) return rc;
  t.aux_in = t.d;
  t.aux_out = t.d;
  if (a->aux_in != nullptr) {
    rc = make_tmap_2d(&t.aux_in, a->aux_in, a->M, a->N, a->ld_aux, cg, BM, 2, rowb);
    if (rc) return rc;
  }
  if (a->aux_out != nullptr) {
    rc = make_tmap_2d(&t.aux_out, a->aux_out, a->M, a->N, a->ld_aux, cg, BM, 2, rowb);
    if (rc) return rc;
  }

  GemmDev d;
  d.M = a->M; d.N = a->N; d.K = a->K;
  d.ldd = a->ldd;
  d.bias = a->bias;
  d.has_aux_out = a->aux_out != nullptr;
  d.mask = a->mask;
  d.drop = make_drop(a->p_drop, a->seed, a->site);
  MH_CHECK(d.drop.thresh == 0 || a->N % 32 == 0, "gemm: a dropout epilogue needs N %% 32 == 0 (one dropout stream word = 32 columns), got N=%d", a->N);
  d.split_k = splits;
  d.delta = a->epilogue == MH_EPI_DELTA ? a->delta : nullptr;
  d.delta_T = a->epilogue == MH_EPI_DELTA ? a->delta_T : 1;
  const int tiles = tiles_m * num_n * splits;
  if (pair) return dispatch_epi_pair(a, t, d, 2 * (tiles < units ? tiles : units), st);
  const int grid = tiles < sms ? tiles : sms;
  if (bn == 256) return dispatch_epi<256>(a, t, d, grid, st);
  return dispatch_epi<128>(a, t, d, grid, st);
}
