// Micro-benchmark: issue/throughput of scalar FFMA/FADD vs packed FFMA2/FADD2/FMUL2 on sm_100a,
// alone and mixed with LDS.64, to size the FFT butterflies of the fused likelihood kernel.
// Prints thread-level FP32 operations per clock per SM (an FMA counts as ONE op here).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int CH = 8; // independent chains per thread

template <int MODE> __global__ void __launch_bounds__(512) bench(float2 *out, int iters, float2 c, float2 d)
{
  __shared__ float2 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x)
    sm[i] = make_float2(1e-3f * i, 1.f);
  __syncthreads();
  float2 a[CH];
#pragma unroll
  for (int k = 0; k < CH; k++)
    a[k] = make_float2(threadIdx.x * 1e-3f + k, k * 0.5f);
  int idx = threadIdx.x;
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int rep = 0; rep < 4; rep++)
    {
#pragma unroll
      for (int k = 0; k < CH; k++)
      {
        if (MODE == 0)
        { // scalar FFMA x2
          a[k].x = fmaf(a[k].x, c.x, d.x);
          a[k].y = fmaf(a[k].y, c.y, d.y);
        }
        else if (MODE == 1) // FFMA2
          a[k] = __ffma2_rn(a[k], c, d);
        else if (MODE == 2)
        { // scalar FADD x2
          a[k].x = a[k].x + d.x;
          a[k].y = a[k].y + d.y;
        }
        else if (MODE == 3) // FADD2
          a[k] = __fadd2_rn(a[k], d);
        else if (MODE == 4) // FADD2 with swap + sign pattern (multiply by i and add)
          a[k] = __fadd2_rn(a[k], make_float2(-a[(k + 1) % CH].y, a[(k + 1) % CH].x));
        else if (MODE == 5)
        { // FFMA2 : LDS.64 = 4 : 1
          a[k] = __ffma2_rn(a[k], c, d);
          if ((k & 3) == 3)
          {
            float2 v = sm[(idx + k) & 1023];
            a[k].x += v.x * 0.f;
          }
        }
        else if (MODE == 6)
        { // alternate FFMA2 (fma pipe) and scalar integer-ish ALU op
          a[k] = __ffma2_rn(a[k], c, d);
          idx = (idx ^ (idx >> 3)) + k;
        }
        else if (MODE == 7)
        { // scalar FFMA x2 + integer ALU op (same issue-slot comparison as mode 6)
          a[k].x = fmaf(a[k].x, c.x, d.x);
          a[k].y = fmaf(a[k].y, c.y, d.y);
          idx = (idx ^ (idx >> 3)) + k;
        }
        else if (MODE == 8)
        { // FADD2 + FFMA2 alternating (do they share a pipe?)
          if (k & 1)
            a[k] = __ffma2_rn(a[k], c, d);
          else
            a[k] = __fadd2_rn(a[k], d);
        }
      }
    }
  }
  float2 s = make_float2((float) idx, 0.f);
#pragma unroll
  for (int k = 0; k < CH; k++)
    s = __fadd2_rn(s, a[k]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> static void run(const char *name, int threads, double flops_per_inner)
{
  int dev = 0;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  const int nsm = prop.multiProcessorCount;
  float2 *out;
  cudaMalloc(&out, sizeof(float2) * nsm * 2 * 1024);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int blocks_per_sm = 1; blocks_per_sm <= 2; blocks_per_sm++)
  {
    bench<MODE><<<nsm * blocks_per_sm, threads>>>(out, 100, make_float2(1.0001f, 0.9999f), make_float2(1e-4f, -1e-4f));
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<MODE><<<nsm * blocks_per_sm, threads>>>(out, iters, make_float2(1.0001f, 0.9999f), make_float2(1e-4f, -1e-4f));
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    const double ops = (double) iters * 4 * CH * flops_per_inner * threads * blocks_per_sm; // per SM
    printf("%-34s threads/SM %4d: %8.3f ms  %7.1f Gop/s/SM  (%.1f ops/clk/SM at the %d MHz nominal clock)\n", name,
           threads * blocks_per_sm, ms, ops / ms / 1e6, ops / (ms * 1e-3) / (clk_khz * 1e3), clk_khz / 1000);
  }
  cudaFree(out);
}

int main()
{
  run<0>("scalar FFMA (2 per step)", 512, 2);
  run<1>("FFMA2", 512, 2);
  run<2>("scalar FADD (2 per step)", 512, 2);
  run<3>("FADD2", 512, 2);
  run<4>("FADD2 swap+NP (x i)", 512, 2);
  run<5>("FFMA2 + LDS.64 every 4th", 512, 2);
  run<6>("FFMA2 + 3 int ALU ops", 512, 2);
  run<7>("2 scalar FFMA + 3 int ALU ops", 512, 2);
  run<8>("FADD2 / FFMA2 alternating", 512, 2);
  return 0;
}
