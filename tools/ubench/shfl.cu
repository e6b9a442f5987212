// Micro-benchmark: warp-shuffle throughput on sm_100a, alone and mixed with LDS.64 / FFMA2 traffic
// (does SHFL share the L1TEX data pipe with shared-memory loads?).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void __launch_bounds__(512) bench(float *out, int iters)
{
  __shared__ float2 sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x)
    sm[i] = make_float2(i * 1e-3f, 1.f);
  __syncthreads();
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; k++)
    a[k] = threadIdx.x + k;
  float2 acc = make_float2(0.f, 0.f);
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int k = 0; k < 8; k++)
    {
      if (MODE == 0 || MODE == 2 || MODE == 3)
        a[k] = __shfl_xor_sync(0xffffffffu, a[k], 1 + (k & 3)) + 1.f;
      if (MODE == 1 || MODE == 2)
      {
        const float2 v = sm[(threadIdx.x + 32 * k + it) & 2047];
        acc.x += v.x;
        acc.y += v.y;
      }
      if (MODE == 3)
        acc = __ffma2_rn(acc, make_float2(1.0001f, 0.9999f), make_float2(a[k], 1.f));
    }
  }
  float s = acc.x + acc.y;
#pragma unroll
  for (int k = 0; k < 8; k++)
    s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + lane;
}

template <int MODE> static void run(const char *name)
{
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount;
  float *out;
  cudaMalloc(&out, sizeof(float) * nsm * 512);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  bench<MODE><<<nsm, 512>>>(out, 100);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  bench<MODE><<<nsm, 512>>>(out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * 1.965e9;
  const double steps = (double) iters * 8 * 16; // warp-level steps per SM (16 warps)
  printf("%-28s %8.3f ms  %.2f cycles per warp-step per SM\n", name, ms, cyc / steps);
  cudaFree(out);
}

int main()
{
  run<0>("SHFL.BFLY + FADD");
  run<1>("LDS.64 + 2 FADD");
  run<2>("SHFL + LDS.64 (both)");
  run<3>("SHFL + FFMA2");
  return 0;
}
