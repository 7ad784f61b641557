// Pipe-throughput microbenchmark for sm_100a (B200): which instruction mixes co-issue?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(512) k(uint32_t* out, uint32_t seed, long long* cyc) {
  uint32_t a[8];
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u; f[i] = __uint_as_float(0x3f000000u | (a[i] & 0x7fffff)); }
  const uint32_t c1 = seed | 1u, c2 = seed * 3u;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 4 || MODE == 5 || MODE == 8 || MODE == 9) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (MODE == 1 || MODE == 4 || MODE == 8) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c1), "r"(c2));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0xD4;" : "+r"(a[i]) : "r"(c2), "r"(c1));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c1), "r"(c2));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0xD4;" : "+r"(a[i]) : "r"(c2), "r"(c1));
      }
      if (MODE == 2 || MODE == 9) {
        uint64_t w;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(a[i]), "r"(0xD2511F53u));
        a[i] = (uint32_t)(w >> 32) ^ (uint32_t)w;   // + 1 LOP
      }
      if (MODE == 3 || MODE == 5 || MODE == 8) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(1.0001f), "f"(0.5f));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(0.9999f), "f"(-0.5f));
      }
      if (MODE == 6) {
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(f[i]), "f"(f[(i + 1) & 7]));
        a[i] ^= r;
      }
      if (MODE == 7) {
        asm volatile("prmt.b32 %0, %0, %1, 0xBB99;" : "+r"(a[i]) : "r"(c1));
        asm volatile("prmt.b32 %0, %0, %1, 0x3120;" : "+r"(a[i]) : "r"(c2));
      }
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i] ^ __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double ops_per_iter_per_thread, int threads) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  k<MODE><<<148, threads>>>(out, 12345u, cyc);
  k<MODE><<<148, threads>>>(out, 12345u, cyc);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("%-34s threads/SM %4d: %8.0f cyc  -> %6.2f cyc per warp-iter-group(8 chains) per SMSP; %6.1f thread-ops/clk/SM\n", name, threads, avg,
         avg / ITERS / (threads / 128.0), ops_per_iter_per_thread * 8 * threads * ITERS / avg);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int th : {128, 256, 512}) {
    run<0>("MUFU.EX2", 1, th);
    run<1>("LOP3 x4", 4, th);
    run<2>("IMAD.WIDE + LOP", 2, th);
    run<3>("FFMA x2", 2, th);
    run<4>("MUFU + LOP3 x4", 5, th);
    run<5>("MUFU + FFMA x2", 3, th);
    run<6>("F2FP + LOP", 2, th);
    run<7>("PRMT x2", 2, th);
    run<8>("MUFU + LOP3 x4 + FFMA x2", 7, th);
    run<9>("MUFU + IMAD.WIDE + LOP", 3, th);
  }
  return 0;
}
