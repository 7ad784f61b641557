// Cost of the generic -> async proxy fence (fence.proxy.async.shared::cta) that every hand-off of register-produced data
// to UMMA / TMA through shared memory needs (attention backward: three per query-block iteration and thread).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fence fence.cu && ./fence
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int ITERS = 1024;

// MODE 0: 4 x st.shared.v4 only   1: + fence.proxy.async   2: + tcgen05.fence::before_thread_sync + __syncwarp
// MODE 3: fence.proxy.async alone (no stores)
template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t* out, long long* cyc) {
  extern __shared__ __align__(16) uint8_t smem[];
  // lane-contiguous 16-byte slots (512 B per warp store: conflict-free), 4 stores 512 B apart, 2 KB per warp
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(smem)) + (threadIdx.x >> 5) * 2048 + (threadIdx.x & 31) * 16;
  uint32_t v = threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (MODE != 3) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + j * 512), "r"(v + it) : "memory");
    }
    if (MODE >= 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (MODE == 2) {
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 512);
  k<MODE><<<148, warps * 32, 64 * 512>>>(out, cyc);
  k<MODE><<<148, warps * 32, 64 * 512>>>(out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < 148; ++i) c += h[i];
  c /= 148;
  printf("%-44s %2d warps: %7.1f clk per iteration  (%s)\n", name, warps, c / ITERS, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 4, 16}) {
    run<0>("4 x st.shared.v4", w);
    run<1>("4 x st.shared.v4 + fence.proxy.async", w);
    run<2>("  + tcgen05.fence::before + syncwarp", w);
    run<3>("fence.proxy.async alone", w);
  }
  return 0;
}
