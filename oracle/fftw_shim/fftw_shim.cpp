// TEST INFRASTRUCTURE — not product code.  See fftw3.h in this directory.
#include "fftw3.h"
#include "../fft_core.hpp"
#include <cstdlib>

struct fftwf_plan_s
{
  int kind; // 0 = c2c, 1 = r2c, 2 = c2r
  int sign;
  offt::Plan2D<float> p;
};

extern "C" {

void *fftwf_malloc(size_t n)
{
  void *p = NULL;
  if (posix_memalign(&p, 64, n ? n : 64) != 0)
    return NULL;
  return p;
}
void fftwf_free(void *p) { free(p); }

static fftwf_plan mk(int kind, int sign, int n0, int n1)
{
  fftwf_plan pl = new fftwf_plan_s;
  pl->kind = kind;
  pl->sign = sign;
  pl->p.init(n0, n1);
  return pl;
}
fftwf_plan fftwf_plan_dft_2d(int n0, int n1, fftwf_complex *, fftwf_complex *,
                             int sign, unsigned)
{
  return mk(0, sign, n0, n1);
}
fftwf_plan fftwf_plan_dft_r2c_2d(int n0, int n1, float *, fftwf_complex *,
                                 unsigned)
{
  return mk(1, -1, n0, n1);
}
fftwf_plan fftwf_plan_dft_c2r_2d(int n0, int n1, fftwf_complex *, float *,
                                 unsigned)
{
  return mk(2, +1, n0, n1);
}
void fftwf_execute_dft_r2c(const fftwf_plan p, float *in, fftwf_complex *out)
{
  offt::r2c_2d<float, float>(p->p, in, (float *) out);
}
void fftwf_execute_dft_c2r(const fftwf_plan p, fftwf_complex *in, float *out)
{
  offt::c2r_2d<float, float>(p->p, (const float *) in, out);
}
void fftwf_execute_dft(const fftwf_plan p, fftwf_complex *in, fftwf_complex *out)
{
  offt::c2c_2d<float, float>(p->p, p->sign, (const float *) in, (float *) out);
}
void fftwf_destroy_plan(fftwf_plan p) { delete p; }
void fftwf_cleanup(void) {}
}
