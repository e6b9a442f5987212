// Micro-benchmark: how much ILP / TLP the packed FP32x2 pipe needs to saturate (FFMA2 and FADD2
// dependent chains, CH independent chains per thread, 1..4 warps per SM sub-partition).
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, int MODE> __global__ void bench(float2 *out, int iters, float2 c, float2 d)
{
  float2 a[CH];
#pragma unroll
  for (int k = 0; k < CH; k++)
    a[k] = make_float2(threadIdx.x * 1e-3f + k, k * 0.5f);
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int rep = 0; rep < 32 / CH; rep++)
#pragma unroll
      for (int k = 0; k < CH; k++)
      {
        if (MODE == 0)
          a[k] = __ffma2_rn(a[k], c, d);
        else if (MODE == 1)
          a[k] = __fadd2_rn(a[k], make_float2(-a[k].y, a[k].x)); // swap + sign pattern, depends on itself
        else
        {
          a[k].x = fmaf(a[k].x, c.x, d.x);
          a[k].y = fmaf(a[k].y, c.y, d.y);
        }
      }
  }
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < CH; k++)
    s = __fadd2_rn(s, a[k]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH, int MODE> static void run(const char *name)
{
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount;
  float2 *out;
  cudaMalloc(&out, sizeof(float2) * nsm * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int threads = 128; threads <= 1024; threads *= 2)
  {
    bench<CH, MODE><<<nsm, threads>>>(out, 100, make_float2(1.0001f, 0.9999f), make_float2(1e-4f, -1e-4f));
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<CH, MODE><<<nsm, threads>>>(out, iters, make_float2(1.0001f, 0.9999f), make_float2(1e-4f, -1e-4f));
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double) iters * 32 * 2 * threads; // FP32 lane-ops per SM
    printf("%-22s chains %d warps/SMSP %d: %6.1f lane-ops/clk/SM (1965 MHz) -> %.2f cycles per dependent op\n", name, CH,
           threads / 128, ops / (ms * 1e-3) / 1.965e9, (ms * 1e-3) * 1.965e9 / ((double) iters * 32 / CH));
  }
  cudaFree(out);
}

int main()
{
  run<1, 0>("FFMA2");
  run<2, 0>("FFMA2");
  run<4, 0>("FFMA2");
  run<1, 1>("FADD2 swap.NP");
  run<2, 1>("FADD2 swap.NP");
  run<1, 2>("2 x scalar FFMA");
  run<2, 2>("2 x scalar FFMA");
  return 0;
}
