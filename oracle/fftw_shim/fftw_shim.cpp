// TEST INFRASTRUCTURE — not product code.  See fftw3.h in this directory.
//
// Two engines behind the nine FFTW symbols the reference uses:
//   * built-in: oracle/fft_core.hpp (double-precision mixed radix; what the golden files were made with), and
//   * Intel oneMKL DFTI, when the environment variable BIOEM_FFT_MKL_LIB names a shared object that exports the DFTI
//     entry points (PyTorch's libtorch_cpu.so is linked against MKL and exports them): the reference arm of bench.py
//     then times the unmodified reference on a vendor FFT of FFTW's class instead of on builder code.  FFTW itself is
//     not in this image.  Only r2c / c2r plans go to MKL; anything that fails falls back to the built-in engine.
#include "fftw3.h"
#include "../fft_core.hpp"
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>

namespace
{
typedef void *DftiHandle;
typedef long MklLong;
// mkl_dfti.h: enum DFTI_CONFIG_PARAM / DFTI_CONFIG_VALUE
enum
{
  DFTI_CONJUGATE_EVEN_STORAGE = 10,
  DFTI_PLACEMENT = 11,
  DFTI_INPUT_STRIDES = 12,
  DFTI_OUTPUT_STRIDES = 13,
  DFTI_PACKED_FORMAT = 21,
  DFTI_THREAD_LIMIT = 27,
  DFTI_REAL = 33,
  DFTI_COMPLEX_COMPLEX = 39,
  DFTI_NOT_INPLACE = 44,
  DFTI_CCE_FORMAT = 57
};
struct Mkl
{
  bool tried = false, ok = false;
  MklLong (*create_s_md)(DftiHandle *, int, MklLong, MklLong *) = nullptr;
  MklLong (*set_value)(DftiHandle, int, ...) = nullptr;
  MklLong (*commit)(DftiHandle) = nullptr;
  MklLong (*forward)(DftiHandle, void *, ...) = nullptr;
  MklLong (*backward)(DftiHandle, void *, ...) = nullptr;
  MklLong (*release)(DftiHandle *) = nullptr;
};
Mkl &mkl()
{
  static Mkl m;
  if (!m.tried)
  {
    m.tried = true;
    const char *path = getenv("BIOEM_FFT_MKL_LIB");
    void *lib = (path && *path) ? dlopen(path, RTLD_LAZY | RTLD_GLOBAL) : nullptr;
    if (lib)
    {
      m.create_s_md = (decltype(m.create_s_md)) dlsym(lib, "DftiCreateDescriptor_s_md");
      m.set_value = (decltype(m.set_value)) dlsym(lib, "DftiSetValue");
      m.commit = (decltype(m.commit)) dlsym(lib, "DftiCommitDescriptor");
      m.forward = (decltype(m.forward)) dlsym(lib, "DftiComputeForward");
      m.backward = (decltype(m.backward)) dlsym(lib, "DftiComputeBackward");
      m.release = (decltype(m.release)) dlsym(lib, "DftiFreeDescriptor");
      m.ok = m.create_s_md && m.set_value && m.commit && m.forward && m.backward && m.release;
    }
    printf("FFT engine: %s\n", m.ok ? "Intel oneMKL DFTI (BIOEM_FFT_MKL_LIB) behind the FFTW-API shim"
                                    : "built-in oracle/fft_core.hpp behind the FFTW-API shim");
    fflush(stdout);
  }
  return m;
}
// real <-> half-spectrum 2-D descriptor, row-major n0 x n1 <-> n0 x (n1/2+1), unnormalised, out of place
DftiHandle mkl_plan(int n0, int n1, bool forward)
{
  Mkl &m = mkl();
  if (!m.ok)
    return nullptr;
  DftiHandle h = nullptr;
  MklLong len[2] = {n0, n1};
  MklLong sr[3] = {0, n1, 1}, sc[3] = {0, n1 / 2 + 1, 1};
  bool good = m.create_s_md(&h, DFTI_REAL, 2, len) == 0 && h;
  good = good && m.set_value(h, DFTI_PLACEMENT, DFTI_NOT_INPLACE) == 0;
  good = good && m.set_value(h, DFTI_CONJUGATE_EVEN_STORAGE, DFTI_COMPLEX_COMPLEX) == 0;
  good = good && m.set_value(h, DFTI_PACKED_FORMAT, DFTI_CCE_FORMAT) == 0;
  good = good && m.set_value(h, DFTI_INPUT_STRIDES, forward ? sr : sc) == 0;
  good = good && m.set_value(h, DFTI_OUTPUT_STRIDES, forward ? sc : sr) == 0;
  if (good)
    m.set_value(h, DFTI_THREAD_LIMIT, (MklLong) 1); // the reference parallelises over images itself
  good = good && m.commit(h) == 0;
  if (!good)
  {
    if (h)
      m.release(&h);
    fprintf(stderr, "fftw_shim: MKL descriptor for %d x %d failed, using the built-in engine\n", n0, n1);
    return nullptr;
  }
  return h;
}
} // namespace

struct fftwf_plan_s
{
  int kind; // 0 = c2c, 1 = r2c, 2 = c2r
  int sign;
  offt::Plan2D<float> p;
  DftiHandle mkl = nullptr;
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
  if (kind == 1 || kind == 2)
    pl->mkl = mkl_plan(n0, n1, kind == 1);
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
  if (p->mkl && mkl().forward(p->mkl, in, out) == 0)
    return;
  offt::r2c_2d<float, float>(p->p, in, (float *) out);
}
void fftwf_execute_dft_c2r(const fftwf_plan p, fftwf_complex *in, float *out)
{
  if (p->mkl && mkl().backward(p->mkl, in, out) == 0)
    return;
  offt::c2r_2d<float, float>(p->p, (const float *) in, out);
}
void fftwf_execute_dft(const fftwf_plan p, fftwf_complex *in, fftwf_complex *out)
{
  offt::c2c_2d<float, float>(p->p, p->sign, (const float *) in, (float *) out);
}
void fftwf_destroy_plan(fftwf_plan p)
{
  if (p && p->mkl)
    mkl().release(&p->mkl);
  delete p;
}
void fftwf_cleanup(void) {}
}
