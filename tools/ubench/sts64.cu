// Micro-benchmark: cost of the fused kernel's 64-bit shared-memory store patterns (exchange tile
// E[k1*ES + c*CS + n2], row slots Y[(j*14+k1)*YS + col]) against a plain aligned pattern.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void __launch_bounds__(256) bench(float *out, int iters)
{
  extern __shared__ float2 sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float2 v = make_float2(lane, warp);
  float2 *base = sm + warp * 1024;
  int off[14];
#pragma unroll
  for (int k = 0; k < 14; k++)
  {
    if (MODE == 0) // aligned rows of 32 float2
      off[k] = k * 32 + lane;
    else if (MODE == 1) // E tile: ES = 34, CS = 17, lane = c*16 + n2
      off[k] = k * 34 + (lane >> 4) * 17 + (lane & 15);
    else if (MODE == 2) // E tile with ES = 36, CS = 18 (16-byte aligned halves)
      off[k] = k * 36 + (lane >> 4) * 18 + (lane & 15);
    else // Y pattern: lane = k1*2 + c, 6 row groups of 14 rows, YS = 114 (CTA-wide array)
      off[k] = ((k % 6) * 14 + (lane >> 1)) * 114 + (lane & 1) + 2 * (k / 6) - warp * 1024 + warp * 2;
  }
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int k = 0; k < 14; k++)
    {
      if (MODE == 3 && lane >= 28)
        continue;
      base[off[k]] = v;
    }
    v.x += 1.f;
    __syncwarp();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x].x;
}

template <int MODE> static void run(const char *name)
{
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int nsm = prop.multiProcessorCount;
  float *out;
  cudaMalloc(&out, sizeof(float) * nsm * 2 * 256);
  const size_t smem = 100 * 1024;
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  bench<MODE><<<nsm * 2, 256, smem>>>(out, 100);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  bench<MODE><<<nsm * 2, 256, smem>>>(out, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * 1.965e9;
  printf("%-40s %.2f cycles per warp-level STS.64 per SM\n", name, cyc / ((double) iters * 14 * 16));
  cudaFree(out);
}

int main()
{
  run<0>("aligned rows (32 float2)");
  run<1>("E tile ES=34 CS=17");
  run<2>("E tile ES=36 CS=18");
  run<3>("Y row slots YS=114 (28 lanes)");
  return 0;
}
