// Data-parallel gradient exchange over NVLink / NVSwitch peer memory (replaces the implicit reduce_add_coalesced /
// broadcast_coalesced of the reference's nn.DataParallel, upstream/melhubert/pretrain_expert.py:28-30).
//
// Why not NCCL for the gradient buckets: every heavy kernel of the step is PERSISTENT -- 74 two-SM GEMM clusters or 148
// attention-backward CTAs that assume the whole chip.  An ncclAllReduce kernel on the side stream occupies whole SMs
// (its CTAs cannot share an SM with a 225 KB / 576-thread GEMM CTA), so a GEMM cluster that cannot be placed waits for a
// full tile wave: the GEMM family fell from 1152 to 845 TFLOP/s as soon as a second GPU joined (SCALE_r01).  The kernels
// here are LIGHT on purpose -- 128 threads, <= 56 registers, no shared memory -- so their CTAs co-reside with the
// persistent CTAs on the same SMs (registers: 608 x 96 + 128 x 56 <= 64 K; shared memory: 225.25 + 1 + 1 KB < 228 KB) and
// take issue slots, not SMs.
//
// Every rank maps every other rank's flat fp32 gradient buffer (CUDA IPC, parallel.PeerGradExchange).  A bucket
// [start, start + count) is cut into `world` 16-byte-aligned shards; rank r owns shard r:
//   reduce-scatter : rank r PULLS shard r of every peer over NVLink and sums in rank order 0 .. world-1 (deterministic,
//                    and bit-identical on all ranks because each element is summed exactly once) into its own buffer;
//   all-gather     : rank r pulls the reduced shard p from rank p, for all p != r.
// Two ways to move the bytes (same sharding, same flags, same results):
//   * copy engines (default, mh_peer_reduce_scatter_ce / mh_peer_all_gather_ce): the pulls are cudaMemcpyAsync calls
//     on peer-mapped pointers -- DMA over NVLink with NO SM involvement -- into a local staging buffer, followed by one
//     light local kernel that adds the staged shards in rank order.  Measured on 2 B200: overlapping SM-driven pulls
//     with the backward saves nothing (each co-resident pull CTA slows the persistent CTA it shares an SM with about
//     as much as running the exchange afterwards costs: +0.9 ms either way), the DMA engines do not have that cost.
//   * SM pulls (mh_peer_reduce_scatter / mh_peer_all_gather): ld.relaxed.sys loops, kept for A/B runs.
// Cross-GPU ordering uses a flag array per rank (peer-mapped too): a kernel first publishes "my inputs are ready" by
// writing its epoch into slot [rank] of every rank's flags (st.release.sys), then waits until all slots of its own
// array reached that epoch (ld.acquire.sys).  Epochs come from a device-resident counter that the last CTA of each
// kernel advances, so CUDA-graph replays stay in step without host involvement; all ranks issue the same kernel
// sequence, hence the same epochs.
#include "mh_b200.h"
#include "mh_common.cuh"

namespace mh {
extern long long g_launches;

constexpr int PEER_MAX = 8;
constexpr int PEER_THREADS = 128;

struct PeerPtrs {
  float* grad[PEER_MAX];                 // flat gradient buffer of rank p (grad[rank] = the local one)
  unsigned long long* flags[PEER_MAX];   // flag array of rank p: PEER_MAX slots
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// peer data: system-scope relaxed loads (never served from this SM's L1, which is not coherent with the peer's writes)
__device__ __forceinline__ float4 ld_peer(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// epoch state of one rank: [0] = epochs completed, [1] = CTA ticket of the kernel in flight
__device__ __forceinline__ unsigned long long peer_enter(const PeerPtrs& pp, unsigned long long* state, int rank, int world) {
  __shared__ unsigned long long s_epoch;
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned long long*>(state) + 1;
  __syncthreads();
  const unsigned long long e = s_epoch;
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(pp.flags[threadIdx.x] + rank, e);
  }
  if (threadIdx.x < world) {
    const unsigned long long* f = pp.flags[rank] + threadIdx.x;
    while (ld_acquire_sys(f) < e) __nanosleep(64);
  }
  __syncthreads();
  return e;
}
__device__ __forceinline__ void peer_exit(unsigned long long* state, unsigned long long e) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(state + 1, 1ull);
    if (t == gridDim.x - 1) {
      state[1] = 0ull;
      __threadfence();
      *reinterpret_cast<volatile unsigned long long*>(state) = e;
    }
  }
}

__device__ __forceinline__ void shard_range(long long start, long long count, int world, int p, long long& lo, long long& hi) {
  const long long per = ((count + world - 1) / world + 3) & ~3LL;  // 16-byte aligned shards (start is 128-byte aligned)
  lo = start + min(per * p, count);
  hi = start + min(per * (p + 1), count);
}

template <int WORLD>
__global__ void __launch_bounds__(PEER_THREADS, 9)  // <= 56 registers: co-resident with a 608-thread attention-backward CTA too
peer_reduce_scatter_kernel(const PeerPtrs pp, unsigned long long* state, long long start, long long count, int rank) {
  const unsigned long long e = peer_enter(pp, state, rank, WORLD);
  long long lo, hi;
  shard_range(start, count, WORLD, rank, lo, hi);
  const long long n4 = (hi - lo) >> 2;  // (count is a multiple of 4: flat slots are 128-byte aligned)
  float* mine = pp.grad[rank] + lo;
  // NVLink pulls are latency-bound (~2000 cycles to a peer's memory): keep WORLD x U 16-byte loads (128 bytes) per
  // thread in flight -- with one CTA per SM that is ~2.4 MB on the wire, enough for the link's ~770 GB/s
  constexpr int U = WORLD <= 2 ? 4 : (WORLD <= 4 ? 2 : 1);
  const long long stride = static_cast<long long>(gridDim.x) * PEER_THREADS;
  long long i = static_cast<long long>(blockIdx.x) * PEER_THREADS + threadIdx.x;
  for (; i + (U - 1) * stride < n4; i += U * stride) {
    float4 v[U][WORLD];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int p = 0; p < WORLD; ++p) v[u][p] = ld_peer(pp.grad[p] + lo + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float4 a = v[u][0];
#pragma unroll
      for (int p = 1; p < WORLD; ++p) { a.x += v[u][p].x; a.y += v[u][p].y; a.z += v[u][p].z; a.w += v[u][p].w; }
      *reinterpret_cast<float4*>(mine + 4 * (i + u * stride)) = a;
    }
  }
  for (; i < n4; i += stride) {
    float4 a = ld_peer(pp.grad[0] + lo + 4 * i);
#pragma unroll
    for (int p = 1; p < WORLD; ++p) {
      const float4 b = ld_peer(pp.grad[p] + lo + 4 * i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    *reinterpret_cast<float4*>(mine + 4 * i) = a;
  }
  peer_exit(state, e);
}

template <int WORLD>
__global__ void __launch_bounds__(PEER_THREADS, 9)
peer_all_gather_kernel(const PeerPtrs pp, unsigned long long* state, long long start, long long count, int rank) {
  const unsigned long long e = peer_enter(pp, state, rank, WORLD);
  float* mine = pp.grad[rank];
#pragma unroll 1
  for (int k = 1; k < WORLD; ++k) {
    const int p = (rank + k) % WORLD;  // every rank starts with a different peer: the pulls spread over the switch
    long long lo, hi;
    shard_range(start, count, WORLD, p, lo, hi);
    const long long n4 = (hi - lo) >> 2;
    const float* src = pp.grad[p] + lo;
    long long i = static_cast<long long>(blockIdx.x) * PEER_THREADS + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * PEER_THREADS;
    for (; i + 7 * stride < n4; i += 8 * stride) {  // 8 independent 16-byte pulls per thread in flight
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = ld_peer(src + 4 * (i + u * stride));
#pragma unroll
      for (int u = 0; u < 8; ++u) *reinterpret_cast<float4*>(mine + lo + 4 * (i + u * stride)) = v[u];
    }
    for (; i < n4; i += stride) *reinterpret_cast<float4*>(mine + lo + 4 * i) = ld_peer(src + 4 * i);
  }
  peer_exit(state, e);
}

// local part of the copy-engine reduce-scatter: mine[i] = sum over ranks p = 0 .. world-1 (ascending, this rank's own
// values at position `rank`) of the staged peer shards; staging slot k holds the shard pulled from rank (k < rank ? k : k + 1)
template <int WORLD>
__global__ void __launch_bounds__(PEER_THREADS, 9)
peer_sum_staged_kernel(float* __restrict__ mine, const float* __restrict__ staging, long long slot_stride, long long n4, int rank) {
  const long long stride = static_cast<long long>(gridDim.x) * PEER_THREADS;
  for (long long i = static_cast<long long>(blockIdx.x) * PEER_THREADS + threadIdx.x; i < n4; i += stride) {
    float4 v[WORLD];
#pragma unroll
    for (int p = 0; p < WORLD; ++p) {
      const float* src = p == rank ? mine : staging + (p < rank ? p : p - 1) * slot_stride;
      v[p] = *reinterpret_cast<const float4*>(src + 4 * i);
    }
    float4 a = v[0];
#pragma unroll
    for (int p = 1; p < WORLD; ++p) { a.x += v[p].x; a.y += v[p].y; a.z += v[p].z; a.w += v[p].w; }
    *reinterpret_cast<float4*>(mine + 4 * i) = a;
  }
}

// barrier only (before the optimizer touches the gradient buffer: every peer has finished pulling from it), with an
// optional small payload: vals[0 .. n) of all ranks are summed (the 2-float (sum of row losses, row count) all-reduce
// that makes the cross-entropy a mean over the GLOBAL masked-frame set).  mailbox: [2][PEER_MAX][16] floats per rank,
// double-buffered by epoch parity (a rank can be at most one exchange ahead of the slowest one).
struct PeerMail {
  float* box[PEER_MAX];
};
__global__ void __launch_bounds__(PEER_THREADS)
peer_barrier_kernel(const PeerPtrs pp, const PeerMail mail, unsigned long long* state, float* vals, int n, int rank, int world) {
  __shared__ unsigned long long s_epoch;
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned long long*>(state) + 1;
  __syncthreads();
  const unsigned long long e = s_epoch;
  const int slot = static_cast<int>(e & 1ull) * PEER_MAX * 16;
  if (n > 0 && threadIdx.x < world * 16) {
    const int p = threadIdx.x >> 4, j = threadIdx.x & 15;
    if (j < n) mail.box[p][slot + rank * 16 + j] = vals[j];
  }
  __syncthreads();
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(pp.flags[threadIdx.x] + rank, e);
    const unsigned long long* f = pp.flags[rank] + threadIdx.x;
    while (ld_acquire_sys(f) < e) __nanosleep(64);
  }
  __syncthreads();
  if (threadIdx.x < n) {
    float s = 0.f;
    for (int p = 0; p < world; ++p)
      s += *reinterpret_cast<volatile float*>(mail.box[rank] + slot + p * 16 + threadIdx.x);
    vals[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    *reinterpret_cast<volatile unsigned long long*>(state) = e;
  }
}
}  // namespace mh

using namespace mh;

static int fill_ptrs(PeerPtrs& pp, void* const* grads, void* const* flags, int world) {
  MH_CHECK(world >= 2 && world <= PEER_MAX, "peer exchange: world size %d not in [2, %d]", world, PEER_MAX);
  for (int p = 0; p < PEER_MAX; ++p) {
    pp.grad[p] = p < world ? static_cast<float*>(grads[p]) : nullptr;
    pp.flags[p] = p < world ? static_cast<unsigned long long*>(flags[p]) : nullptr;
    MH_CHECK(p >= world || (pp.grad[p] != nullptr && pp.flags[p] != nullptr), "peer exchange: null peer pointer (rank %d)", p);
  }
  return 0;
}

template <int W>
static void launch_rs_ag(int which, const PeerPtrs& pp, unsigned long long* state, long long start, long long count, int rank,
                         int ctas, cudaStream_t st) {
  if (which == 0)
    peer_reduce_scatter_kernel<W><<<ctas, PEER_THREADS, 0, st>>>(pp, state, start, count, rank);
  else
    peer_all_gather_kernel<W><<<ctas, PEER_THREADS, 0, st>>>(pp, state, start, count, rank);
}

static int peer_rs_ag(int which, void* const* grads, void* const* flags, void* state, long long start, long long count,
                      int rank, int world, int ctas, void* stream) {
  PeerPtrs pp;
  if (int rc = fill_ptrs(pp, grads, flags, world)) return rc;
  MH_CHECK(rank >= 0 && rank < world && state != nullptr, "peer exchange: bad rank %d / state", rank);
  MH_CHECK(start % 4 == 0 && count % 4 == 0 && count > 0, "peer exchange: bucket [%lld, +%lld) must be 16-byte aligned", start, count);
  if (ctas <= 0) ctas = sm_count();  // one light CTA per SM
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unsigned long long* s = static_cast<unsigned long long*>(state);
  switch (world) {
    case 2: launch_rs_ag<2>(which, pp, s, start, count, rank, ctas, st); break;
    case 3: launch_rs_ag<3>(which, pp, s, start, count, rank, ctas, st); break;
    case 4: launch_rs_ag<4>(which, pp, s, start, count, rank, ctas, st); break;
    case 5: launch_rs_ag<5>(which, pp, s, start, count, rank, ctas, st); break;
    case 6: launch_rs_ag<6>(which, pp, s, start, count, rank, ctas, st); break;
    case 7: launch_rs_ag<7>(which, pp, s, start, count, rank, ctas, st); break;
    default: launch_rs_ag<8>(which, pp, s, start, count, rank, ctas, st); break;
  }
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

extern "C" int mh_peer_reduce_scatter(void* const* grads, void* const* flags, void* state, long long start, long long count,
                                      int rank, int world, int ctas, void* stream) {
  return peer_rs_ag(0, grads, flags, state, start, count, rank, world, ctas, stream);
}

extern "C" int mh_peer_all_gather(void* const* grads, void* const* flags, void* state, long long start, long long count,
                                  int rank, int world, int ctas, void* stream) {
  return peer_rs_ag(1, grads, flags, state, start, count, rank, world, ctas, stream);
}

extern "C" int mh_peer_barrier_sum(void* const* grads, void* const* flags, void* const* mailboxes, void* state, float* vals,
                                   int n, int rank, int world, void* stream) {
  PeerPtrs pp;
  if (int rc = fill_ptrs(pp, grads, flags, world)) return rc;
  MH_CHECK(n >= 0 && n <= 16 && (n == 0 || vals != nullptr), "peer barrier: payload of %d floats (max 16)", n);
  PeerMail mail;
  for (int p = 0; p < PEER_MAX; ++p) mail.box[p] = (p < world && mailboxes != nullptr) ? static_cast<float*>(mailboxes[p]) : nullptr;
  MH_CHECK(n == 0 || mailboxes != nullptr, "peer barrier: a payload needs the mailboxes");
  peer_barrier_kernel<<<1, PEER_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      pp, mail, static_cast<unsigned long long*>(state), vals, n, rank, world);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}

// ---- copy-engine variants: barrier kernel (1 light CTA) + cudaMemcpyAsync pulls over peer-mapped pointers (+ local sum)
static void host_shard(long long start, long long count, int world, int p, long long& lo, long long& hi) {
  const long long per = ((count + world - 1) / world + 3) & ~3LL;
  lo = start + (per * p < count ? per * p : count);
  hi = start + (per * (p + 1) < count ? per * (p + 1) : count);
}

template <int W>
static void launch_sum(float* mine, const float* staging, long long slot_stride, long long n4, int rank, cudaStream_t st) {
  long long g = (n4 + PEER_THREADS - 1) / PEER_THREADS;
  if (g > sm_count()) g = sm_count();
  if (g < 1) g = 1;
  peer_sum_staged_kernel<W><<<static_cast<int>(g), PEER_THREADS, 0, st>>>(mine, staging, slot_stride, n4, rank);
}

extern "C" int mh_peer_reduce_scatter_ce(void* const* grads, void* const* flags, void* state, float* staging,
                                         long long staging_elems, long long start, long long count, int rank, int world,
                                         void* stream) {
  PeerPtrs pp;
  if (int rc = fill_ptrs(pp, grads, flags, world)) return rc;
  MH_CHECK(rank >= 0 && rank < world && state != nullptr && staging != nullptr, "peer exchange: bad rank %d / state / staging", rank);
  MH_CHECK(start % 4 == 0 && count % 4 == 0 && count > 0, "peer exchange: bucket [%lld, +%lld) must be 16-byte aligned", start, count);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long lo, hi;
  host_shard(start, count, world, rank, lo, hi);
  const long long n = hi - lo;
  const long long slot = (n + 31) & ~31LL;
  MH_CHECK(slot * (world - 1) <= staging_elems, "peer exchange: staging buffer too small (%lld < %lld floats)", staging_elems, slot * (world - 1));
  PeerMail mail = {};
  peer_barrier_kernel<<<1, PEER_THREADS, 0, st>>>(pp, mail, static_cast<unsigned long long*>(state), nullptr, 0, rank, world);
  MH_LAUNCH_CHECK();
  ++g_launches;
  if (n > 0) {
    for (int k = 1; k < world; ++k) {  // start with the next rank: the pulls of all ranks spread over the switch
      const int p = (rank + k) % world;
      const int s = p < rank ? p : p - 1;
      MH_CUDA(cudaMemcpyAsync(staging + s * slot, pp.grad[p] + lo, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
    }
    float* mine = pp.grad[rank] + lo;
    switch (world) {
      case 2: launch_sum<2>(mine, staging, slot, n >> 2, rank, st); break;
      case 3: launch_sum<3>(mine, staging, slot, n >> 2, rank, st); break;
      case 4: launch_sum<4>(mine, staging, slot, n >> 2, rank, st); break;
      case 5: launch_sum<5>(mine, staging, slot, n >> 2, rank, st); break;
      case 6: launch_sum<6>(mine, staging, slot, n >> 2, rank, st); break;
      case 7: launch_sum<7>(mine, staging, slot, n >> 2, rank, st); break;
      default: launch_sum<8>(mine, staging, slot, n >> 2, rank, st); break;
    }
    MH_LAUNCH_CHECK();
    ++g_launches;
  }
  return 0;
}

extern "C" int mh_peer_all_gather_ce(void* const* grads, void* const* flags, void* state, long long start, long long count,
                                     int rank, int world, void* stream) {
  PeerPtrs pp;
  if (int rc = fill_ptrs(pp, grads, flags, world)) return rc;
  MH_CHECK(rank >= 0 && rank < world && state != nullptr, "peer exchange: bad rank %d / state", rank);
  MH_CHECK(start % 4 == 0 && count % 4 == 0 && count > 0, "peer exchange: bucket [%lld, +%lld) must be 16-byte aligned", start, count);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  PeerMail mail = {};
  peer_barrier_kernel<<<1, PEER_THREADS, 0, st>>>(pp, mail, static_cast<unsigned long long*>(state), nullptr, 0, rank, world);
  MH_LAUNCH_CHECK();
  ++g_launches;
  for (int k = 1; k < world; ++k) {
    const int p = (rank + k) % world;
    long long lo, hi;
    host_shard(start, count, world, p, lo, hi);
    if (hi > lo)
      MH_CUDA(cudaMemcpyAsync(pp.grad[rank] + lo, pp.grad[p] + lo, sizeof(float) * (hi - lo), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}
