// Host-side one-off input preparation behind the C ABI (include/bioem_b200.h):
// the quantities the reference derives once before its main loop.  Pure CPU code,
// float/double arithmetic with the reference's promotion rules (myfloat_t = float).
#include "../../include/bioem_b200.h"
#include <cmath>
#include <cstring>
#include <vector>

extern "C" {

// reference param.cpp:601-607: defocus [micro-m] -> CTF phase; the prior centre and
// width of the defocus are scaled by the same factor.
void bioem_b200_host_defocus_to_phase(float startDefocus, float endDefocus, float elecwavel, float *startPhase,
                                      float *endPhase, float *Priordefcent, float *sigmaPriordefo)
{
  const double k = M_PI * 2.f * 10000 * elecwavel; // ((pi*2)*10000)*lambda, in double
  // the reference multiplies left to right starting from the float defocus
  *startPhase = (float) (startDefocus * M_PI * 2.f * 10000 * elecwavel);
  *endPhase = (float) (endDefocus * M_PI * 2.f * 10000 * elecwavel);
  if (Priordefcent)
    *Priordefcent = (float) (*Priordefcent * k);
  if (sigmaPriordefo)
    *sigmaPriordefo = (float) (*sigmaPriordefo * k);
}

// reference param.cpp:1336-1620.  CTF mode fills the half-spectrum directly; the
// reference writes row i and its "mirror" N-i-1 for i = 0..N/2 in that order, later
// writes overwriting earlier ones — reproduced by resolving, for
// every output row, which i wrote it last.
int bioem_b200_host_ctf_table(int N, float pixelSize, int usepsf, float startAmp, float endAmp, int nAmp,
                              float startPhase, float endPhase, int nPhase, float startEnv, float endEnv,
                              int nEnv, float *refCTF, float *CtfParam4, float *grids)
{
  if (N <= 0 || nAmp <= 0 || nPhase <= 0 || nEnv <= 0)
    return -1;
  float stepAmp = (endAmp - startAmp) / (float) nAmp;
  float stepPhase = (endPhase - startPhase) / (float) nPhase;
  float stepEnv = (endEnv - startEnv) / (float) nEnv;
  // single-point grids: the step variable takes the start value (it feeds volu; quirk Q2)
  if (nAmp == 1)
    stepAmp = startAmp;
  if (nPhase == 1)
    stepPhase = startPhase;
  if (nEnv == 1)
    stepEnv = startEnv;
  if (grids)
  {
    grids[0] = stepAmp;
    grids[1] = stepPhase;
    grids[2] = stepEnv;
  }
  const int nTot = nAmp * nPhase * nEnv;
  if (!refCTF && !CtfParam4)
    return nTot;
  if (usepsf)
    return -2; // point-spread functions are built in real space: bioem_b200_host_psf_kernels + upload_ctf_real
  const int nc = N / 2 + 1;
  const size_t F = (size_t) N * nc;
  std::vector<int> writer(N, -1);
  for (int i = 0; i < nc; i++)
  {
    writer[i] = i;
    writer[N - i - 1] = i;
  }
  std::vector<float> row(nc);
  int n = 0;
  for (int ia = 0; ia < nAmp; ia++)
  {
    const float amp = (float) ia * stepAmp + startAmp;
    for (int ip = 0; ip < nPhase; ip++)
    {
      const float phase = (float) ip * stepPhase + startPhase;
      for (int ie = 0; ie < nEnv; ie++, n++)
      {
        const float env = (float) ie * stepEnv + startEnv;
        if (CtfParam4)
        {
          CtfParam4[4 * n + 0] = amp;
          CtfParam4[4 * n + 1] = phase;
          CtfParam4[4 * n + 2] = env;
          CtfParam4[4 * n + 3] = 0.f;
        }
        if (!refCTF)
          continue;
        float *cur = refCTF + (size_t) n * F * 2;
        memset(cur, 0, sizeof(float) * F * 2);
        // value at (0,0) normalises the kernel
        float norm = 0.f;
        for (int i = 0; i < nc; i++)
        {
          for (int j = 0; j < nc; j++)
          {
            const float radsq = (float) (i * i + j * j) / N / N / pixelSize / pixelSize;
            const float ctf = (float) (exp(-env * radsq / 2.) * (-amp * cos(phase * radsq / 2.) -
                                                                 sqrtf(1 - amp * amp) * sin(phase * radsq / 2.)));
            if (i == 0 && j == 0)
              norm = ctf;
            row[j] = ctf / norm;
          }
          for (int r = 0; r < N; r++)
            if (writer[r] == i)
              for (int j = 0; j < nc; j++)
                cur[2 * ((size_t) r * nc + j)] = row[j];
        }
      }
    }
  }
  return nTot;
}

// reference param.cpp:1336-1536, USE_PSF: the kernels are defined in REAL space (radially symmetric
// around pixel (0,0), periodic), normalised to unit sum; the reference then takes their r2c FFT
// (param.cpp:1521) -- here that transform runs on the device (bioem_b200_upload_ctf_real).
int bioem_b200_host_psf_kernels(int N, float pixelSize, float startAmp, float endAmp, int nAmp, float startPhase,
                                float endPhase, int nPhase, float startEnv, float endEnv, int nEnv, float *kernels,
                                float *CtfParam4, float *grids)
{
  if (N <= 0 || nAmp <= 0 || nPhase <= 0 || nEnv <= 0)
    return -1;
  float stepAmp = (endAmp - startAmp) / (float) nAmp;
  float stepPhase = (endPhase - startPhase) / (float) nPhase;
  float stepEnv = (endEnv - startEnv) / (float) nEnv;
  if (nAmp == 1)
    stepAmp = startAmp;
  if (nPhase == 1)
    stepPhase = startPhase;
  if (nEnv == 1)
    stepEnv = startEnv;
  if (grids)
  {
    grids[0] = stepAmp;
    grids[1] = stepPhase;
    grids[2] = stepEnv;
  }
  const int nTot = nAmp * nPhase * nEnv;
  if (!kernels && !CtfParam4)
    return nTot;
  // widest envelope must fit the kernel length (param.cpp:1403-1409)
  if (sqrt(1. / ((float) nEnv * stepEnv + startEnv)) > float(N) / 2.0)
    return -3;
  const int nctfmax = N / 2;
  int n = 0;
  for (int ia = 0; ia < nAmp; ia++)
  {
    const float amp = (float) ia * stepAmp + startAmp;
    for (int ip = 0; ip < nPhase; ip++)
    {
      const float phase = (float) ip * stepPhase + startPhase;
      for (int ie = 0; ie < nEnv; ie++, n++)
      {
        const float env = (float) ie * stepEnv + startEnv;
        if (CtfParam4)
        {
          CtfParam4[4 * n + 0] = amp;
          CtfParam4[4 * n + 1] = phase;
          CtfParam4[4 * n + 2] = env;
          CtfParam4[4 * n + 3] = 0.f;
        }
        if (!kernels)
          continue;
        float *cur = kernels + (size_t) n * N * N;
        float normctf = 0.f;
        for (int i = 0; i < N; i++)
          for (int j = 0; j < N; j++)
          {
            const int ri = (i < nctfmax + 1) ? i : N - i;
            const int rj = (j < nctfmax + 1) ? j : N - j;
            const float radsq = (float) (ri * ri + rj * rj) * pixelSize * pixelSize;
            const float ctf = (float) (exp(-radsq * env / 2.0) * (-amp * cos(radsq * phase / 2.0) -
                                                                  sqrtf(1 - amp * amp) * sin(radsq * phase / 2.0)));
            cur[(size_t) i * N + j] = ctf;
            normctf += ctf;
          }
        for (size_t k = 0; k < (size_t) N * N; k++)
          cur[k] = cur[k] / normctf;
      }
    }
  }
  return nTot;
}

// reference param.cpp:1600-1607
float bioem_b200_host_volu(float voluang, int GridSpaceCenter, float pixelSize, int maxDisplaceCenter, int nAmp,
                           float gridEnvelop, float gridCTF_phase, float sigmaPriorbctf, float sigmaPriordefo,
                           float sigmaPrioramp)
{
  const float G = (float) GridSpaceCenter;
  // float products up to the first double term, double afterwards, narrowed at the end
  const float head = voluang * G * pixelSize * G * pixelSize;
  double v = head / ((2.f * (float) maxDisplaceCenter + 1.));
  v = v / (2.f * (float) (maxDisplaceCenter + 1.));
  v = v / (float) nAmp * gridEnvelop * gridCTF_phase / 4.f / M_PI / sqrt(2.f * M_PI) / sigmaPriorbctf /
      sigmaPriordefo / sigmaPrioramp;
  return (float) v;
}

// reference model.cpp (NormDen = running float sum of densities) and :604-672
float bioem_b200_host_model_prepare(bioem_b200_model_point *pts, int A, int center)
{
  float normDen = 0.f;
  for (int n = 0; n < A; n++)
    normDen += pts[n].density;
  if (center)
  {
    float cm[3] = {0.f, 0.f, 0.f};
    for (int n = 0; n < A; n++)
      for (int k = 0; k < 3; k++)
        cm[k] += pts[n].pos[k] * pts[n].density;
    for (int k = 0; k < 3; k++)
      cm[k] /= normDen;
    for (int n = 0; n < A; n++)
      for (int k = 0; k < 3; k++)
        pts[n].pos[k] -= cm[k];
  }
  return normDen;
}

// reference map.cpp:811-845: statistics accumulated in file order (transposed with
// respect to memory), then img = img/std - mean/std
void bioem_b200_host_normalise_map(float *img, int N)
{
  float st = 0.f, st2 = 0.f;
  for (int j = 0; j < N; j++)
    for (int i = 0; i < N; i++)
    {
      const float c = img[(size_t) i * N + j];
      st += c;
      st2 += c * c;
    }
  st /= float(N * N);
  st2 = sqrtf(st2 / float(N * N) - st * st);
  for (size_t k = 0; k < (size_t) N * N; k++)
    img[k] = img[k] / st2 - st / st2;
}

// reference bioem.cpp:1144-1149
double bioem_b200_host_final_logprob(const bioem_b200_config *cfg, double Total, double Constoadd)
{
  return log(Total) + Constoadd + 0.5 * log(M_PI) + (1 - cfg->Ntotpi * 0.5) * (log(2 * M_PI) + 1) +
         log(cfg->volu);
}
}
