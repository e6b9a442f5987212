// TEST INFRASTRUCTURE — not product code.
//
// CPU restatement ("oracle") of the BioEM likelihood hot path, stage by stage,
// following the reference sources cited at each function (paths relative to
// /root/reference).  It is the checker for the CUDA path: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library.  The product (libbioem_b200.so) never links or calls it.
//
// Parity pinning: this restatement is itself checked against the UNMODIFIED
// reference built from /root/reference (oracle/_ref/bioEM_ref, see Makefile) in
// tests/test_oracle.py, both on the final Output_Probabilities and on
// the per-evaluation -DDEBUG_PROB stream (bioem_algorithm.h:88-128), and against
// the golden outputs of that binary committed under tests/golden/.
//
// Arithmetic: myfloat_t = float, myprob_t = double (include/defs.h:48-66).  The
// file is compiled with -ffp-contract=off and without -ffast-math, i.e. it
// evaluates every expression in source order with the C++ promotion rules the
// reference's expressions have; the reference binary itself is built with
// -ffast-math (CMakeLists.txt:33), so it differs from this by rounding noise
// (SURVEY quirk Q7).
//
// FFTs: FFTW 3 (single precision) is a third-party dependency absent from this
// image; its documented semantics are restated in fft_core.hpp.  oracle_set_fft
// selects float (what fftwf computes in) or double internals.
#include "fft_core.hpp"
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef float myfloat_t;
typedef double myprob_t;

extern "C" {

// Mirror of the fields of bioem_param_device (include/param.h:26-47) plus the few
// host-side parameters the path needs (pixelSize, shiftX/Y, doquater).
struct OracleCfg
{
  int NumberPixels;
  int maxDisplaceCenter;
  int GridSpaceCenter;
  int writeAngles;
  int tousepsf;
  int doquater;
  int shiftX;
  int shiftY;
  float pixelSize;
  float Ntotpi;
  float volu;
  float sigmaPriorbctf;
  float sigmaPriordefo;
  float Priordefcent;
  float sigmaPrioramp;
  float Priorampcent;
};

// include/map.h:116-129 (40 bytes) and :131-135 (16 bytes)
struct OracleProbMap
{
  double Total;
  double Constoadd;
  int max_prob_cent_x, max_prob_cent_y, max_prob_orient, max_prob_conv;
  float max_prob_norm, max_prob_mu;
};
struct OracleProbAngle
{
  double forAngles;
  double ConstAngle;
};

static int g_fft_double = 0;
void oracle_set_fft(int use_double) { g_fft_double = use_double; }

} // extern "C"

namespace
{
template <typename T> const offt::Plan2D<T> &plan_for(int n)
{
  static std::mutex mu;
  static std::map<int, std::unique_ptr<offt::Plan2D<T>>> plans;
  std::lock_guard<std::mutex> lk(mu);
  auto it = plans.find(n);
  if (it == plans.end())
  {
    std::unique_ptr<offt::Plan2D<T>> p(new offt::Plan2D<T>);
    p->init(n, n);
    it = plans.emplace(n, std::move(p)).first;
  }
  return *it->second;
}
void fft_r2c(int n, const float *in, float *out)
{
  if (g_fft_double)
    offt::r2c_2d<double, float>(plan_for<double>(n), in, out);
  else
    offt::r2c_2d<float, float>(plan_for<float>(n), in, out);
}
void fft_c2r(int n, const float *in, float *out)
{
  if (g_fft_double)
    offt::c2r_2d<double, float>(plan_for<double>(n), in, out);
  else
    offt::c2r_2d<float, float>(plan_for<float>(n), in, out);
}
} // namespace

extern "C" {

// ---------------------------------------------------------------------------
// param.cpp:601-607 — defocus [micro-m] -> phase, and the same scaling applied
// to the defocus prior centre / width (CTF mode only).
// ---------------------------------------------------------------------------
void oracle_defocus_to_phase(float startDefocus, float endDefocus, float elecwavel,
                             float *startPhase, float *endPhase, float *Priordefcent,
                             float *sigmaPriordefo)
{
  *startPhase = startDefocus * M_PI * 2.f * 10000 * elecwavel;
  *endPhase = endDefocus * M_PI * 2.f * 10000 * elecwavel;
  *Priordefcent *= M_PI * 2.f * 10000 * elecwavel;
  *sigmaPriordefo *= M_PI * 2.f * 10000 * elecwavel;
}

// ---------------------------------------------------------------------------
// param.cpp:1336-1620 CalculateRefCTF — kernel table, CtfParam list, volu.
//   refCTF   [C][F][2] float, CtfParam [C][3] (amp, phase, env)
//   grids[3] out: gridCTF_amp, gridCTF_phase, gridEnvelop (incl. the n==1 quirk Q2)
// Returns C.
// ---------------------------------------------------------------------------
int oracle_ctf_table(int N, float pixelSize, int usepsf, float startAmp, float endAmp,
                     int nAmp, float startPhase, float endPhase, int nPhase,
                     float startEnv, float endEnv, int nEnv, float *refCTF,
                     float *CtfParam, float *grids)
{
  const int nc = N / 2 + 1;
  const size_t F = (size_t) N * nc;
  const int nctfmax = N / 2;
  myfloat_t gridCTF_amp = (endAmp - startAmp) / (myfloat_t) nAmp;
  myfloat_t gridCTF_phase = (endPhase - startPhase) / (myfloat_t) nPhase;
  myfloat_t gridEnvelop = (endEnv - startEnv) / (myfloat_t) nEnv;
  if (nAmp == 1)
    gridCTF_amp = startAmp; // param.cpp:1373-1376
  if (nPhase == 1)
    gridCTF_phase = startPhase;
  if (nEnv == 1)
    gridEnvelop = startEnv;
  if (grids)
  {
    grids[0] = gridCTF_amp;
    grids[1] = gridCTF_phase;
    grids[2] = gridEnvelop;
  }
  if (!refCTF)
    return nAmp * nPhase * nEnv;

  std::vector<float> localCTF((size_t) N * N);
  int n = 0;
  for (int iamp = 0; iamp < nAmp; iamp++)
  {
    myfloat_t amp = (myfloat_t) iamp * gridCTF_amp + startAmp;
    for (int iphase = 0; iphase < nPhase; iphase++)
    {
      myfloat_t phase = (myfloat_t) iphase * gridCTF_phase + startPhase;
      for (int ienv = 0; ienv < nEnv; ienv++)
      {
        myfloat_t env = (myfloat_t) ienv * gridEnvelop + startEnv;
        float *cur = &refCTF[(size_t) n * F * 2];
        memset(cur, 0, sizeof(float) * F * 2);
        myfloat_t normctf = 0.0;
        myfloat_t radsq, ctf;
        if (usepsf)
        {
          // param.cpp:1466-1535: real-space kernel, normalised to unit sum, r2c
          for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++)
            {
              int ri = (i < nctfmax + 1) ? i : N - i;
              int rj = (j < nctfmax + 1) ? j : N - j;
              radsq = (myfloat_t)(ri * ri + rj * rj) * pixelSize * pixelSize;
              ctf = exp(-radsq * env / 2.0) *
                    (-amp * cos(radsq * phase / 2.0) -
                     sqrtf((1 - amp * amp)) * sin(radsq * phase / 2.0));
              localCTF[(size_t) i * N + j] = (myfloat_t) ctf;
              normctf += localCTF[(size_t) i * N + j];
            }
          for (size_t i = 0; i < (size_t) N * N; i++)
            localCTF[i] = localCTF[i] / normctf;
          fft_r2c(N, localCTF.data(), cur);
        }
        else
        {
          // param.cpp:1548-1570: directly in Fourier space; the mirrored write goes
          // to row N-i-1 (quirk Q1) and later rows overwrite earlier ones.
          for (int i = 0; i < nc; i++)
            for (int j = 0; j < nc; j++)
            {
              radsq = (myfloat_t)(i * i + j * j) / N / N / pixelSize / pixelSize;
              ctf = exp(-env * radsq / 2.) *
                    (-amp * cos(phase * radsq / 2.) -
                     sqrtf((1 - amp * amp)) * sin(phase * radsq / 2.));
              if (i == 0 && j == 0)
                normctf = (myfloat_t) ctf;
              cur[2 * ((size_t) i * nc + j)] = ctf / normctf;
              cur[2 * ((size_t) i * nc + j) + 1] = 0;
              cur[2 * ((size_t)(N - i - 1) * nc + j)] = ctf / normctf;
              cur[2 * ((size_t)(N - i - 1) * nc + j) + 1] = 0;
            }
        }
        CtfParam[3 * n + 0] = amp;
        CtfParam[3 * n + 1] = phase;
        CtfParam[3 * n + 2] = env;
        n++;
      }
    }
  }
  return n;
}

// param.cpp:1600-1607 (volu) with voluang from param.cpp:1131,1324 (lists:
// 1/O * priorMod) — the caller passes voluang.
float oracle_volu(float voluang, int GridSpaceCenter, float pixelSize, int maxDisplaceCenter,
                  int numberGridPointsCTF_amp, float gridEnvelop, float gridCTF_phase,
                  float sigmaPriorbctf, float sigmaPriordefo, float sigmaPrioramp)
{
  myfloat_t volu =
      voluang * (myfloat_t) GridSpaceCenter * pixelSize * (myfloat_t) GridSpaceCenter *
      pixelSize / ((2.f * (myfloat_t) maxDisplaceCenter + 1.)) /
      (2.f * (myfloat_t)(maxDisplaceCenter + 1.)) / (myfloat_t) numberGridPointsCTF_amp *
      gridEnvelop * gridCTF_phase / 4.f / M_PI / sqrt(2.f * M_PI) / sigmaPriorbctf /
      sigmaPriordefo / sigmaPrioramp;
  return volu;
}
float oracle_voluang_list(int nOrient, float priorMod)
{
  myfloat_t voluang = 1. / (myfloat_t) nOrient * priorMod;
  return voluang;
}

// ---------------------------------------------------------------------------
// model.cpp:419-601 (NormDen = sequential float sum of densities) and
// model.cpp:604-672 centerDensityMass (sequential variant).
// pts: [A][5] = x y z radius density, modified in place.  Returns NormDen.
// ---------------------------------------------------------------------------
float oracle_model_prepare(float *pts, int A, int center)
{
  myfloat_t NormDen = 0.0;
  for (int n = 0; n < A; n++)
    NormDen += pts[5 * n + 4];
  if (center)
  {
    myfloat_t r[3] = {0.f, 0.f, 0.f};
    for (int n = 0; n < A; n++)
      for (int k = 0; k < 3; k++)
        r[k] += pts[5 * n + k] * pts[5 * n + 4];
    for (int k = 0; k < 3; k++)
      r[k] /= NormDen;
    for (int n = 0; n < A; n++)
      for (int k = 0; k < 3; k++)
        pts[5 * n + k] -= r[k];
  }
  return NormDen;
}

// map.cpp:830-845: per-image zero-mean / unit-std normalisation done by the MRC
// reader (float accumulators, file order).  img is N*N in memory order.
void oracle_normalise_map(float *img, int N)
{
  myfloat_t st = 0.0, st2 = 0.0;
  // the reader accumulates in file order (j outer, i inner) = transposed memory order
  for (int j = 0; j < N; j++)
    for (int i = 0; i < N; i++)
    {
      float c = img[(size_t) i * N + j];
      st += c;
      st2 += c * c;
    }
  st /= float(N * N);
  st2 = sqrtf(st2 / float(N * N) - st * st);
  for (size_t k = 0; k < (size_t) N * N; k++)
    img[k] = img[k] / st2 - st / st2;
}

// ---------------------------------------------------------------------------
// map.cpp:603-630 RefMap.precalculate + bioem.cpp:2087-2107 calcross_cor +
// map.cpp:557-601 PreCalculateMapsFFT.
// ---------------------------------------------------------------------------
void oracle_particle_prepare(const float *maps, int M, int N, float *RefMapsFFT,
                             float *sum_RefMap, float *sumsquare_RefMap)
{
  const size_t F = (size_t) N * (N / 2 + 1);
#pragma omp parallel for
  for (int m = 0; m < M; m++)
  {
    const float *mp = &maps[(size_t) m * N * N];
    myfloat_t sum = 0.0, sumsquare = 0.0;
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++)
      {
        sum += mp[(size_t) i * N + j];
        sumsquare += mp[(size_t) i * N + j] * mp[(size_t) i * N + j];
      }
    sum_RefMap[m] = sum;
    sumsquare_RefMap[m] = sumsquare;
    fft_r2c(N, mp, &RefMapsFFT[(size_t) m * F * 2]);
  }
}

// ---------------------------------------------------------------------------
// bioem.cpp:1604-1853 createProjection.  angle = (pos[0],pos[1],pos[2],quat4).
// pts [A][5]; projReal (optional, N*N) receives the scaled real-space image;
// mapFFT [F][2].  Returns the number of points skipped as out-of-frame.
// ---------------------------------------------------------------------------
int oracle_projection(const OracleCfg *cfg, const float *pts, int A, float NormDen,
                      const float *angle, float *mapFFT, float *projReal)
{
  const int N = cfg->NumberPixels;
  const myfloat_t pixelSize = cfg->pixelSize;
  std::vector<float> localproj((size_t) N * N, 0.f);
  myfloat_t rotmat[3][3];
  if (cfg->doquater)
  {
    myfloat_t quater[4] = {angle[0], angle[1], angle[2], angle[3]};
    rotmat[0][0] = 1 - 2 * quater[1] * quater[1] - 2 * quater[2] * quater[2];
    rotmat[1][0] = 2 * (quater[0] * quater[1] - quater[2] * quater[3]);
    rotmat[2][0] = 2 * (quater[0] * quater[2] + quater[1] * quater[3]);
    rotmat[0][1] = 2 * (quater[0] * quater[1] + quater[2] * quater[3]);
    rotmat[1][1] = 1 - 2 * quater[0] * quater[0] - 2 * quater[2] * quater[2];
    rotmat[2][1] = 2 * (quater[1] * quater[2] - quater[0] * quater[3]);
    rotmat[0][2] = 2 * (quater[0] * quater[2] - quater[1] * quater[3]);
    rotmat[1][2] = 2 * (quater[1] * quater[2] + quater[0] * quater[3]);
    rotmat[2][2] = 1 - 2 * quater[0] * quater[0] - 2 * quater[1] * quater[1];
  }
  else
  {
    myfloat_t alpha = angle[0], beta = angle[1], gam = angle[2];
    rotmat[0][0] = cosf(gam) * cosf(alpha) - cosf(beta) * sinf(alpha) * sinf(gam);
    rotmat[0][1] = cosf(gam) * sinf(alpha) + cosf(beta) * cosf(alpha) * sinf(gam);
    rotmat[0][2] = sinf(gam) * sinf(beta);
    rotmat[1][0] = -sinf(gam) * cosf(alpha) - cosf(beta) * sinf(alpha) * cosf(gam);
    rotmat[1][1] = -sinf(gam) * sinf(alpha) + cosf(beta) * cosf(alpha) * cosf(gam);
    rotmat[1][2] = cosf(gam) * sinf(beta);
    rotmat[2][0] = sinf(beta) * sinf(alpha);
    rotmat[2][1] = -sinf(beta) * cosf(alpha);
    rotmat[2][2] = cosf(beta);
  }
  int skipped = 0;
  myfloat_t tempden = 0.0;
  for (int n = 0; n < A; n++)
  {
    myfloat_t rp[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < 3; k++)
      for (int j = 0; j < 3; j++)
        rp[k] += rotmat[k][j] * pts[5 * n + j];
    const myfloat_t radius = pts[5 * n + 3], density = pts[5 * n + 4];
    int i, j;
    if (radius <= pixelSize)
    {
      i = floorf(rp[0] / pixelSize + (myfloat_t) N / 2.0f + 0.5f);
      j = floorf(rp[1] / pixelSize + (myfloat_t) N / 2.0f + 0.5f);
      if (i < 0 || j < 0 || i >= N || j >= N)
        skipped++;
      else
      {
        localproj[(size_t) i * N + j] += density;
        tempden += density;
      }
    }
    else
    {
      i = floorf(rp[0] / pixelSize + (myfloat_t) N / 2.0f + 0.5f) - cfg->shiftX;
      j = floorf(rp[1] / pixelSize + (myfloat_t) N / 2.0f + 0.5f) - cfg->shiftY;
      const int irad = int(radius / pixelSize) + 1;
      const myfloat_t rad2 = radius * radius;
      if (i < irad || j < irad || i >= N - irad || j >= N - irad)
        skipped++;
      else
      {
        for (int ii = i - irad; ii < i + irad + 1; ii++)
          for (int jj = j - irad; jj < j + irad + 1; jj++)
          {
            myfloat_t dist = ((myfloat_t)(ii - i) * (ii - i) + (jj - j) * (jj - j)) *
                             pixelSize * pixelSize;
            if (dist < rad2)
            {
              // float numerator / double denominator, accumulated into floats
              // (bioem.cpp:1792-1798)
              localproj[(size_t) ii * N + jj] += pixelSize * pixelSize * 2 *
                                                 sqrtf(rad2 - dist) * density * 3 /
                                                 (4 * M_PI * radius * rad2);
              tempden += pixelSize * pixelSize * 2 * sqrtf(rad2 - dist) * density * 3 /
                         (4 * M_PI * radius * rad2);
            }
          }
      }
    }
  }
  const myfloat_t ratioDen = NormDen / tempden;
  for (size_t k = 0; k < (size_t) N * N; k++)
    localproj[k] *= ratioDen;
  if (projReal)
    memcpy(projReal, localproj.data(), sizeof(float) * (size_t) N * N);
  fft_r2c(N, localproj.data(), mapFFT);
  return skipped;
}

// ---------------------------------------------------------------------------
// bioem.cpp:1855-1923 createConvolutedProjectionMap.
// ---------------------------------------------------------------------------
void oracle_convolve(int N, const float *lproj, const float *refCTF, float *localmultFFT,
                     float *sumC_out, float *sumsquareC_out)
{
  const int nc = N / 2 + 1;
  for (int i = 0; i < N * nc; i++)
  {
    localmultFFT[2 * i] = (lproj[2 * i] * refCTF[2 * i] + lproj[2 * i + 1] * refCTF[2 * i + 1]);
    localmultFFT[2 * i + 1] =
        (lproj[2 * i + 1] * refCTF[2 * i] - lproj[2 * i] * refCTF[2 * i + 1]);
  }
  myfloat_t sumC = localmultFFT[0];
  myfloat_t sumsquareC = 0;
  int jloopend = nc;
  if ((N & 1) == 0)
    jloopend--;
  for (int i = 0; i < N; i++)
  {
    for (int j = 1; j < jloopend; j++)
    {
      int k = i * nc + j;
      sumsquareC += (localmultFFT[2 * k] * localmultFFT[2 * k] +
                     localmultFFT[2 * k + 1] * localmultFFT[2 * k + 1]) *
                    2;
    }
    int k = i * nc;
    sumsquareC +=
        localmultFFT[2 * k] * localmultFFT[2 * k] + localmultFFT[2 * k + 1] * localmultFFT[2 * k + 1];
    if ((N & 1) == 0)
    {
      k += nc - 1;
      sumsquareC += localmultFFT[2 * k] * localmultFFT[2 * k] +
                    localmultFFT[2 * k + 1] * localmultFFT[2 * k + 1];
    }
  }
  myfloat_t norm2 = (myfloat_t)(N * N);
  sumsquareC = sumsquareC / norm2;
  *sumC_out = sumC;
  *sumsquareC_out = sumsquareC;
}

// ---------------------------------------------------------------------------
// bioem.cpp:1435-1459 calculateCCFFT: lCC = c2r( conv * conj(RefMapFFT) ).
// ---------------------------------------------------------------------------
void oracle_cross_correlation(int N, const float *localConvFFT, const float *RefMapFFT,
                              float *lCC)
{
  const int nc = N / 2 + 1;
  std::vector<float> localCCT((size_t) N * nc * 2);
  for (int i = 0; i < N * nc; i++)
  {
    localCCT[2 * i] =
        localConvFFT[2 * i] * RefMapFFT[2 * i] + localConvFFT[2 * i + 1] * RefMapFFT[2 * i + 1];
    localCCT[2 * i + 1] =
        localConvFFT[2 * i + 1] * RefMapFFT[2 * i] - localConvFFT[2 * i] * RefMapFFT[2 * i + 1];
  }
  fft_c2r(N, localCCT.data(), lCC);
}

// ---------------------------------------------------------------------------
// bioem_algorithm.h:18-70 calc_logpro.
// ---------------------------------------------------------------------------
double oracle_calc_logpro(const OracleCfg *param, float amp, float pha, float env, float sum,
                          float sumsquare, float crossproMapConv, float sumref,
                          float sumsquareref)
{
  const myfloat_t Ntotpi = param->Ntotpi;
  const myprob_t ForLogProb = sumsquare * Ntotpi - sum * sum;
  const myprob_t firstele =
      Ntotpi * (sumsquareref * sumsquare - crossproMapConv * crossproMapConv) +
      2 * sumref * sum * crossproMapConv - sumsquareref * sum * sum - sumref * sumref * sumsquare;
  myprob_t logpro =
      (3 - Ntotpi) * 0.5 * log(firstele) + (Ntotpi * 0.5 - 2) * log((Ntotpi - 2) * ForLogProb);
  if (!param->tousepsf)
  {
    logpro -= env * env / 2. / param->sigmaPriorbctf / param->sigmaPriorbctf -
              (pha - param->Priordefcent) * (pha - param->Priordefcent) / 2. /
                  param->sigmaPriordefo / param->sigmaPriordefo -
              (amp - param->Priorampcent) * (amp - param->Priorampcent) / 2. /
                  param->sigmaPrioramp / param->sigmaPrioramp;
  }
  else
  {
    myprob_t envF, phaF;
    envF = 4. * M_PI * M_PI * env / (env * env + pha * pha);
    phaF = 4. * M_PI * M_PI * pha / (env * env + pha * pha);
    logpro -= envF * envF / 2. / param->sigmaPriorbctf / param->sigmaPriorbctf -
              (phaF - param->Priordefcent) * (phaF - param->Priordefcent) / 2. /
                  param->sigmaPriordefo / param->sigmaPriordefo -
              (amp - param->Priorampcent) * (amp - param->Priorampcent) / 2. /
                  param->sigmaPrioramp / param->sigmaPrioramp;
  }
  return logpro;
}

} // extern "C"

namespace
{
// Optional per-evaluation trace (the analogue of the reference's -DDEBUG_PROB
// stream): float-narrowed logpro for every (o, c, displacement) of ONE image.
struct Trace
{
  int image = -1;
  float *logpro = nullptr; // [O][C][D]
  float *value = nullptr;  // [O][C][D]
  size_t D = 0, C = 0;
};

// bioem_algorithm.h:72-142 calProb
inline void calProb(const OracleCfg *param, int iRefMap, int iOrient, int iConv, float amp,
                    float pha, float env, float sumC, float sumsquareC, float value, int disx,
                    int disy, float sumRef, float sumsquareRef, OracleProbMap &pProbMap,
                    OracleProbAngle *pProbAngle, double *second, float *trace_lp,
                    float *trace_val)
{
  myfloat_t logpro =
      oracle_calc_logpro(param, amp, pha, env, sumC, sumsquareC, value, sumRef, sumsquareRef);
  if (trace_lp)
  {
    *trace_lp = logpro;
    *trace_val = value;
  }
  if (pProbMap.Constoadd < logpro)
  {
    if (second)
      *second = pProbMap.Constoadd; // the dethroned maximum becomes the runner-up
    pProbMap.Total *= exp(-logpro + pProbMap.Constoadd);
    pProbMap.Constoadd = logpro;
    pProbMap.max_prob_cent_x = -disx;
    pProbMap.max_prob_cent_y = -disy;
    pProbMap.max_prob_orient = iOrient;
    pProbMap.max_prob_conv = iConv;
    pProbMap.max_prob_norm =
        -(-sumC * sumRef + param->Ntotpi * value) / (sumC * sumC - sumsquareC * param->Ntotpi);
    pProbMap.max_prob_mu =
        -(-sumC * value + sumsquareC * sumRef) / (sumC * sumC - sumsquareC * param->Ntotpi);
  }
  else if (second && *second < logpro)
    *second = logpro;
  pProbMap.Total += exp(logpro - pProbMap.Constoadd);
  if (param->writeAngles)
  {
    if (pProbAngle->ConstAngle < logpro)
    {
      pProbAngle->forAngles *= exp(-logpro + pProbAngle->ConstAngle);
      pProbAngle->ConstAngle = logpro;
    }
    pProbAngle->forAngles += exp(logpro - pProbAngle->ConstAngle);
  }
}

// bioem_algorithm.h:144-198 doRefMapFFT — the exact enumeration order.
inline void doRefMapFFT(const OracleCfg *param, int iRefMap, int iOrient, int iConv, float amp,
                        float pha, float env, float sumC, float sumsquareC, const float *lCC,
                        float sumRef, float sumsquareRef, OracleProbMap &pm, OracleProbAngle *pa,
                        double *second, float *trace_lp, float *trace_val)
{
  const int N = param->NumberPixels, maxD = param->maxDisplaceCenter, G = param->GridSpaceCenter;
  size_t t = 0;
#define ORACLE_EVAL(cx, cy, dx, dy)                                                              \
  calProb(param, iRefMap, iOrient, iConv, amp, pha, env, sumC, sumsquareC,                       \
          (myfloat_t) lCC[(cx) * N + (cy)] / (myfloat_t)(N * N), dx, dy, sumRef, sumsquareRef,   \
          pm, pa, second, trace_lp ? trace_lp + t : nullptr, trace_val ? trace_val + t : nullptr); \
  t++;
  for (int cent_x = 0; cent_x <= maxD; cent_x = cent_x + G)
  {
    for (int cent_y = 0; cent_y <= maxD; cent_y = cent_y + G)
    {
      ORACLE_EVAL(cent_x, cent_y, cent_x, cent_y)
    }
    for (int cent_y = N - maxD; cent_y < N; cent_y = cent_y + G)
    {
      ORACLE_EVAL(cent_x, cent_y, cent_x, cent_y - N)
    }
  }
  for (int cent_x = N - maxD; cent_x < N; cent_x = cent_x + G)
  {
    for (int cent_y = 0; cent_y <= maxD; cent_y = cent_y + G)
    {
      ORACLE_EVAL(cent_x, cent_y, cent_x - N, cent_y)
    }
    for (int cent_y = N - maxD; cent_y < N; cent_y = cent_y + G)
    {
      ORACLE_EVAL(cent_x, cent_y, cent_x - N, cent_y - N)
    }
  }
#undef ORACLE_EVAL
}
} // namespace

extern "C" {

// Number of displacements doRefMapFFT enumerates per likelihood (Algo 1; quirk Q3).
int oracle_num_displacements(int N, int maxD, int G)
{
  int a = 0;
  for (int c = 0; c <= maxD; c += G)
    a++;
  for (int c = N - maxD; c < N; c += G)
    a++;
  return a * a;
}

// ---------------------------------------------------------------------------
// bioem.cpp:659-903 run() main loop with Algo 1 (bioem.cpp:1390-1408):
// for o in [oBegin, oEnd): projection; for c: convolution; for m: cross
// correlation + doRefMapFFT.  State initialised as bioem.cpp:681-699.
//   pts [A][5], angles [O][4], refCTF [C][F][2], CtfParam [C][3],
//   RefMapsFFT [M][F][2], sumRef/sumsqRef [M]
//   probMap [M], probAngle [O][M] (layout angle*nMaps+map, map.h:147-150) or NULL
//   second [M] (optional): runner-up float logpro per image (for near-tie tests)
//   trace_image >= 0: trace_lp/trace_val [O][C][D] for that image
// ---------------------------------------------------------------------------
int oracle_run(const OracleCfg *cfg, const float *pts, int A, float NormDen, const float *angles,
               int O, int oBegin, int oEnd, const float *refCTF, const float *CtfParam, int C,
               const float *RefMapsFFT, const float *sumRef, const float *sumsqRef, int M,
               OracleProbMap *probMap, OracleProbAngle *probAngle, double *second,
               int trace_image, float *trace_lp, float *trace_val)
{
  const int N = cfg->NumberPixels;
  const size_t F = (size_t) N * (N / 2 + 1);
  const size_t D = (size_t) oracle_num_displacements(N, cfg->maxDisplaceCenter, cfg->GridSpaceCenter);
  const double MIN_PROB = -999999.;
  for (int m = 0; m < M; m++)
  {
    probMap[m].Total = 0.0;
    probMap[m].Constoadd = MIN_PROB;
    probMap[m].max_prob_cent_x = probMap[m].max_prob_cent_y = 0;
    probMap[m].max_prob_orient = probMap[m].max_prob_conv = 0;
    probMap[m].max_prob_norm = probMap[m].max_prob_mu = 0.f;
    if (second)
      second[m] = MIN_PROB;
    if (cfg->writeAngles && probAngle)
      for (int o = 0; o < O; o++)
      {
        probAngle[(size_t) o * M + m].forAngles = 0.0;
        probAngle[(size_t) o * M + m].ConstAngle = MIN_PROB;
      }
  }
  std::vector<float> proj(F * 2), conv(F * 2);
  OracleProbAngle dummy;
  for (int o = oBegin; o < oEnd; o++)
  {
    oracle_projection(cfg, pts, A, NormDen, &angles[4 * o], proj.data(), nullptr);
    for (int c = 0; c < C; c++)
    {
      float sumC, sumsquareC;
      oracle_convolve(N, proj.data(), &refCTF[(size_t) c * F * 2], conv.data(), &sumC, &sumsquareC);
      const float amp = CtfParam[3 * c], pha = CtfParam[3 * c + 1], env = CtfParam[3 * c + 2];
#pragma omp parallel
      {
        std::vector<float> lCC((size_t) N * N);
#pragma omp for schedule(dynamic, 1)
        for (int m = 0; m < M; m++)
        {
          oracle_cross_correlation(N, conv.data(), &RefMapsFFT[(size_t) m * F * 2], lCC.data());
          OracleProbAngle *pa = (cfg->writeAngles && probAngle) ? &probAngle[(size_t) o * M + m] : &dummy;
          OracleCfg local = *cfg;
          if (!probAngle)
            local.writeAngles = 0;
          const bool tr = (m == trace_image) && trace_lp;
          const size_t toff = ((size_t)(o - oBegin) * C + c) * D;
          doRefMapFFT(&local, m, o, c, amp, pha, env, sumC, sumsquareC, lCC.data(), sumRef[m],
                      sumsqRef[m], probMap[m], pa, second ? &second[m] : nullptr,
                      tr ? trace_lp + toff : nullptr, tr ? trace_val + toff : nullptr);
        }
      }
    }
  }
  return 0;
}

// bioem.cpp:1144-1149: final log-posterior of an image.
double oracle_final_logprob(const OracleCfg *cfg, double Total, double Constoadd)
{
  return log(Total) + Constoadd + 0.5 * log(M_PI) +
         (1 - cfg->Ntotpi * 0.5) * (log(2 * M_PI) + 1) + log(cfg->volu);
}

// Number of worker threads the oracle will use (for bench.py's cpu_baseline.cores).
int oracle_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// Plain forward / inverse transforms, exposed so tests can pin the FFT restatement
// against numpy.
void oracle_fft_r2c(int N, const float *in, float *out) { fft_r2c(N, in, out); }
void oracle_fft_c2r(int N, const float *in, float *out) { fft_c2r(N, in, out); }

} // extern "C"
