// Direct-DFT path for image edges the fused FFT kernel is not instantiated for (odd NUMBER_PIXELS, prime factors
// above 7, edges above 512 -- the reference accepts any edge, param.cpp:140-152, and handles odd ones in its
// Parseval weights, bioem.cpp:1893-1918).  Same stages, same arithmetic in the epilogue, but every transform is a
// plain O(N^2)-per-line DFT against a twiddle table in shared memory: about 25x the fused kernel's time per
// likelihood at N = 225 -- a path that gives the right answer on every edge, not a fast one.
//
// Spectra live in the reference's own layout here: [k0][k1], k1 <= N/2 contiguous (NC = N/2 + 1 complex per row);
// consecutive maps are S = 2 * map4 float2 apart (N * NC rounded up to whole float4).
//
//   gen_dft_rows_kernel / gen_dft_cols_kernel   forward r2c 2-D transform           (bioem.cpp:1848, map.cpp:585)
//   gen_conv_kernel                             createConvolutedProjectionMap       (bioem.cpp:1855-1923)
//   gen_corr_kernel                             calculateCCFFT on the window only   (bioem.cpp:1435-1459)
//   gen_fold_kernel                             doRefMapFFT + calc_logpro + calProb (bioem_algorithm.h:18-198)
#pragma once
#include "bioem_kernels.cuh"

namespace bioem
{

// forward DFT along the contiguous axis of real images: scratch[img][r][k1] = sum_c img[r][c] exp(-2 pi i k1 c / N)
// (tw[j] = exp(+2 pi i j / N); tempden / normDen: the projection's density normalisation as in fft_rows_kernel)
__global__ void __launch_bounds__(256) gen_dft_rows_kernel(const float *__restrict__ imgs, const double *__restrict__ tempden,
                                                           int nbands, float normDen, const float2 *__restrict__ tw, int N,
                                                           float2 *__restrict__ scratch)
{
  extern __shared__ float2 g_sm[];
  float2 *TW = g_sm;                                 // [N]
  float *row = reinterpret_cast<float *>(g_sm + N); // [N]
  const int NC = N / 2 + 1;
  const int img = blockIdx.y, r = blockIdx.x, tid = threadIdx.x;
  float scale = 1.f;
  if (tempden)
  {
    double t = 0.0;
    for (int b = 0; b < nbands; b++)
      t += tempden[(size_t) img * nbands + b];
    scale = __fdiv_rn(normDen, (float) t);
  }
  const float *src = imgs + ((size_t) img * N + r) * N;
  for (int i = tid; i < N; i += blockDim.x)
  {
    TW[i] = tw[i];
    row[i] = __fmul_rn(src[i], scale);
  }
  __syncthreads();
  for (int k = tid; k < NC; k += blockDim.x)
  {
    float re = 0.f, im = 0.f;
    int idx = 0;
    for (int c = 0; c < N; c++)
    {
      const float2 w = TW[idx];
      re = fmaf(row[c], w.x, re);
      im = fmaf(-row[c], w.y, im);
      idx += k;
      if (idx >= N)
        idx -= N;
    }
    scratch[((size_t) img * N + r) * NC + k] = make_float2(re, im);
  }
}

// forward DFT along the other axis: out[img][k0][k1] = sum_r scratch[img][r][k1] exp(-2 pi i k0 r / N)
__global__ void __launch_bounds__(256) gen_dft_cols_kernel(const float2 *__restrict__ scratch, const float2 *__restrict__ tw, int N,
                                                           size_t S, float2 *__restrict__ out)
{
  extern __shared__ float2 g_sm[];
  float2 *TW = g_sm; // [N]
  const int NC = N / 2 + 1;
  const int img = blockIdx.y, k0 = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < N; i += blockDim.x)
    TW[i] = tw[i];
  __syncthreads();
  const float2 *src = scratch + (size_t) img * N * NC;
  for (int k1 = tid; k1 < NC; k1 += blockDim.x)
  {
    float re = 0.f, im = 0.f;
    int idx = 0;
    for (int r = 0; r < N; r++)
    {
      const float2 w = TW[idx]; // conj: exp(-i..)
      const float2 a = src[(size_t) r * NC + k1];
      re += a.x * w.x + a.y * w.y;
      im += a.y * w.x - a.x * w.y;
      idx += k0;
      if (idx >= N)
        idx -= N;
    }
    out[(size_t) img * S + (size_t) k0 * NC + k1] = make_float2(re, im);
  }
}

// stage 2 on the reference layout: V = P * conj(K_c), sumC, sumsquareC (Parseval weights bioem.cpp:1893-1918: 1 for
// k1 = 0 and, when N is even, k1 = N/2; 2 otherwise) and the displacement-independent term.  Grid as ctf_conv_kernel.
__global__ void __launch_bounds__(256) gen_conv_kernel(const float2 *__restrict__ proj, const float2 *__restrict__ ctf,
                                                       const double *__restrict__ prior, float2 *__restrict__ conv,
                                                       ConvParam *__restrict__ cpar, int C, int N, size_t S, float Ntotpi)
{
  const int NC = N / 2 + 1, F = N * NC;
  const int tid = threadIdx.x, c = blockIdx.x, ob = blockIdx.y;
  const size_t oslot = (size_t) ob * C + c;
  const float2 *P = proj + (size_t) ob * S;
  const float2 *K = ctf + (size_t) c * S;
  float2 *V = conv + oslot * S;
  float acc = 0.f, sumC = 0.f;
  for (int i = tid; i < F; i += blockDim.x)
  {
    const float2 a = P[i], k = K[i];
    float2 v;
    v.x = a.x * k.x + a.y * k.y;
    v.y = a.y * k.x - a.x * k.y;
    V[i] = v;
    const int k1 = i % NC;
    const float w = (k1 == 0 || 2 * k1 == N) ? 1.f : 2.f;
    acc += w * (v.x * v.x + v.y * v.y);
    if (i == 0)
      sumC = v.x;
  }
  __shared__ float red[256];
  __shared__ float s_sumC;
  red[tid] = acc;
  if (tid == 0)
    s_sumC = sumC;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1)
  {
    if (tid < s)
      red[tid] += red[tid + s];
    __syncthreads();
  }
  if (tid == 0)
  {
    const float ssC = __fdiv_rn(red[0], (float) (N * N));
    ConvParam cp;
    cp.sumC = s_sumC;
    cp.sumsqC = ssC;
    const float fl = __fsub_rn(__fmul_rn(ssC, Ntotpi), __fmul_rn(s_sumC, s_sumC));
    cp.Bterm = ((double) Ntotpi * 0.5 - 2.0) * log((double) __fsub_rn(Ntotpi, 2.f) * (double) fl) - prior[c];
    cpar[oslot] = cp;
  }
}

// Correlation window of one (conv spectrum, particle) item: values[item][wx*nw + wy] = lCC[x][y] / N^2 for the
// window displacements x = wl[wx], y = wl[wy] (doRefMapFFT's enumeration, bioem_algorithm.h:156-197), where lCC is
// the c2r inverse transform of conv * conj(particle).  Two partial DFTs through shared memory:
//   T[wx][k1] = sum_k0 Z[k0][k1] exp(+2 pi i k0 x / N)                       (thread = k1 x a block of XB window rows)
//   lCC[x][y] = sum_k1 w(k1) Re(T[wx][k1] exp(+2 pi i k1 y / N)),  w = 1 for the self-conjugate columns, else 2
// (a c2r transform sees exactly that much of the half-spectrum: quirk Q1 needs no special case here).
// Grid: x = particle, y = conv spectrum of the batch.
constexpr int GEN_XB = 16;
__global__ void __launch_bounds__(256) gen_corr_kernel(const float2 *__restrict__ convs, const float2 *__restrict__ refs,
                                                       const float2 *__restrict__ tw, const int *__restrict__ wl, int N, size_t S,
                                                       int nw, int M, float invNN, float *__restrict__ values)
{
  extern __shared__ float2 g_sm[];
  const int NC = N / 2 + 1;
  float2 *TW = g_sm;                                    // [N]
  float2 *T = g_sm + N;                                 // [nw][NC]
  int *WL = reinterpret_cast<int *>(T + (size_t) nw * NC); // [nw]
  const int tid = threadIdx.x, m = blockIdx.x, oc = blockIdx.y;
  for (int i = tid; i < N; i += blockDim.x)
    TW[i] = tw[i];
  for (int i = tid; i < nw; i += blockDim.x)
    WL[i] = wl[i];
  __syncthreads();
  const float2 *V = convs + (size_t) oc * S;
  const float2 *R = refs + (size_t) m * S;
  // work item = (block of XB window rows, column k1): consecutive threads take consecutive columns of the same block
  // (coalesced operand loads, broadcast twiddle reads)
  const int nxb = (nw + GEN_XB - 1) / GEN_XB;
  for (int item = tid; item < nxb * NC; item += blockDim.x)
  {
    const int k1 = item % NC, w0 = (item / NC) * GEN_XB;
    {
      float2 acc[GEN_XB];
      int idx[GEN_XB], step[GEN_XB];
#pragma unroll
      for (int j = 0; j < GEN_XB; j++)
      {
        acc[j] = make_float2(0.f, 0.f);
        idx[j] = 0;
        step[j] = w0 + j < nw ? WL[w0 + j] : 0;
      }
      for (int k0 = 0; k0 < N; k0++)
      {
        const float2 v = V[(size_t) k0 * NC + k1], r = R[(size_t) k0 * NC + k1];
        const float2 z = bfft::cmulc(v, r); // conv * conj(particle)
#pragma unroll
        for (int j = 0; j < GEN_XB; j++)
        {
          const float2 w = TW[idx[j]];
          acc[j].x += z.x * w.x - z.y * w.y;
          acc[j].y += z.x * w.y + z.y * w.x;
          idx[j] += step[j];
          if (idx[j] >= N)
            idx[j] -= N;
        }
      }
#pragma unroll
      for (int j = 0; j < GEN_XB; j++)
        if (w0 + j < nw)
          T[(size_t) (w0 + j) * NC + k1] = acc[j];
    }
  }
  __syncthreads();
  float *out = values + ((size_t) oc * M + m) * nw * nw;
  for (int i = tid; i < nw * nw; i += blockDim.x)
  {
    const int wx = i / nw, wy = i % nw;
    const int y = WL[wy];
    const float2 *Tr = T + (size_t) wx * NC;
    float s = 0.f;
    int idx = 0;
    for (int k1 = 0; k1 < NC; k1++)
    {
      const float2 t = Tr[k1], w = TW[idx];
      const float term = t.x * w.x - t.y * w.y;
      s += (k1 == 0 || 2 * k1 == N) ? term : 2.f * term;
      idx += y;
      if (idx >= N)
        idx -= N;
    }
    out[i] = s * invNN;
  }
}

// calc_logpro + calProb for the items of one batch, one CTA per particle, items in (orientation, CTF) order: for
// every displacement firstele in FP32 in the reference's operation order (as the fused kernel's epilogue), logpro =
// (float)(a log(firstele) + B) in double narrowed to float (bioem_algorithm.h:84), the first maximum in enumeration
// order keeps the record (:96), Total / Constoadd accumulate in double exactly as calProb does, likelihood by
// likelihood; the per-orientation ANG_PROB rows are the same fold restricted to one orientation.
__global__ void __launch_bounds__(128) gen_fold_kernel(const float *__restrict__ values, const ConvParam *__restrict__ cpar,
                                                       const float *__restrict__ sumRef, const float *__restrict__ sumsqRef,
                                                       int OBcur, int C, int M, int nw, int o_base, float Nt, double acoef,
                                                       Running *__restrict__ state, ProbAngleOut *__restrict__ angles)
{
  const int m = blockIdx.x, tid = threadIdx.x, n = nw * nw;
  __shared__ float s_f[128];
  __shared__ int s_i[128];
  __shared__ double s_d[128];
  const float sR = sumRef[m], ssR = sumsqRef[m];
  Running run = state[m]; // (only thread 0's copy is used)
  for (int ob = 0; ob < OBcur; ob++)
  {
    double anConst = kMinProb, anTotal = 0.0;
    for (int c = 0; c < C; c++)
    {
      const int oc = ob * C + c;
      const ConvParam cp = cpar[oc];
      const float *v = values + ((size_t) oc * M + m) * n;
      const float f_a = __fmul_rn(ssR, cp.sumsqC);
      const float f_b = __fmul_rn(__fmul_rn(2.f, sR), cp.sumC);
      const float f_c = __fmul_rn(__fmul_rn(ssR, cp.sumC), cp.sumC);
      const float f_d = __fmul_rn(__fmul_rn(sR, sR), cp.sumsqC);
      auto logpro = [&](float raw) {
        float f = __fmul_rn(raw, raw);
        f = __fadd_rn(f_a, -f);
        f = __fmaf_rn(Nt, f, __fmul_rn(f_b, raw));
        f = __fadd_rn(f, -f_c);
        f = __fadd_rn(f, -f_d);
        return (float) __fma_rn(acoef, log((double) f), cp.Bterm);
      };
      float best = __int_as_float(0xff800000);
      int besti = 0x7fffffff;
      for (int i = tid; i < n; i += 128)
      {
        const float lp = logpro(v[i]);
        if (lp > best) // strict: the first maximum of this thread's (ascending) indices
        {
          best = lp;
          besti = i;
        }
      }
      s_f[tid] = best;
      s_i[tid] = besti;
      __syncthreads();
      for (int s = 64; s > 0; s >>= 1)
      {
        if (tid < s)
        {
          const float of = s_f[tid + s];
          const int oi = s_i[tid + s];
          if (of > s_f[tid] || (of == s_f[tid] && oi < s_i[tid]))
          {
            s_f[tid] = of;
            s_i[tid] = oi;
          }
        }
        __syncthreads();
      }
      const float lpmax = s_f[0];
      const int imax = s_i[0];
      double e = 0.0;
      for (int i = tid; i < n; i += 128)
        e += exp((double) logpro(v[i]) - (double) lpmax);
      s_d[tid] = e;
      __syncthreads();
      for (int s = 64; s > 0; s >>= 1)
      {
        if (tid < s)
          s_d[tid] += s_d[tid + s];
        __syncthreads();
      }
      if (tid == 0)
      {
        e = s_d[0];
        if (run.Const < (double) lpmax)
        {
          run.Total = run.Total * exp(run.Const - (double) lpmax) + e;
          run.Const = (double) lpmax;
          run.lpf = lpmax;
          run.orient = o_base + ob;
          run.conv = c;
          run.lin = imax;
          run.v = v[imax];
          run.sumC = cp.sumC;
          run.sumsqC = cp.sumsqC;
        }
        else
          run.Total += e * exp((double) lpmax - run.Const);
        if (anConst < (double) lpmax)
        {
          anTotal = anTotal * exp(anConst - (double) lpmax) + e;
          anConst = (double) lpmax;
        }
        else
          anTotal += e * exp((double) lpmax - anConst);
      }
      __syncthreads();
    }
    if (tid == 0 && angles)
    {
      ProbAngleOut a;
      a.forAngles = anTotal;
      a.ConstAngle = anConst;
      angles[(size_t) (o_base + ob) * M + m] = a;
    }
  }
  if (tid == 0)
    state[m] = run;
}

} // namespace bioem
