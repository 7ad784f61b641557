// TMEM read bandwidth of one SM (sm_100a): how many bytes per clock do tcgen05.ld instructions deliver to the register
// file, as a function of the number of reading warps and of the load shape?  The attention backward reads 160 KB of fp32
// accumulators (S, dP, dQ) per 128 x 128 query / key block; this number decides what bounds it.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu && ./tmem_ld
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(slot));
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

#define LD32(r, addr)                                                                                                     \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19," \
               "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                                                 \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),   \
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),       \
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),      \
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                   \
               : "r"(addr))
#define LD8(r, addr)                                                                                      \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                    \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) \
               : "r"(addr))

constexpr int ITERS = 512;

// SHAPE 32: x32 loads (4 KB per warp instruction), 8: x8 loads (1 KB); DEPTH loads in flight before each wait::ld
template <int SHAPE, int DEPTH>
__global__ void __launch_bounds__(512) k(uint32_t* out, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (SHAPE == 32) {
      uint32_t r[DEPTH][32];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) LD32(r[d], base + ((it + d) & 7) * 32 + (warp >> 2) * 0);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) acc ^= r[d][0] ^ r[d][31];
    } else {
      uint32_t r[DEPTH * 4][8];
#pragma unroll
      for (int d = 0; d < DEPTH * 4; ++d) LD8(r[d], base + ((it * 4 + d) & 31) * 8);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int d = 0; d < DEPTH * 4; ++d) acc ^= r[d][0] ^ r[d][7];
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

template <int SHAPE, int DEPTH>
void run(int warps) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  k<SHAPE, DEPTH><<<148, warps * 32>>>(out, cyc);
  k<SHAPE, DEPTH><<<148, warps * 32>>>(out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < 148; ++i) c += h[i];
  c /= 148;
  const double bytes = static_cast<double>(warps) * ITERS * DEPTH * 4096.0;
  printf("x%-2d  depth %d  %2d warps: %8.0f clk  %6.1f B/clk/SM  (%s)\n", SHAPE, DEPTH, warps, c, bytes / c, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8, 16}) run<32, 1>(w);
  for (int w : {4, 8, 16}) run<32, 2>(w);
  for (int w : {4, 8, 16}) run<8, 1>(w);
  for (int w : {4, 16}) run<8, 2>(w);
  return 0;
}
