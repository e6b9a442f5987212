/* TEST INFRASTRUCTURE — not product code.
 *
 * Minimal stand-in for <fftw3.h> so that the UNMODIFIED reference sources under
 * /root/reference compile in an image without FFTW 3.  It declares exactly the
 * nine single-precision symbols the reference uses (nm -u of its objects):
 *   fftwf_malloc fftwf_free fftwf_plan_dft_2d fftwf_plan_dft_r2c_2d
 *   fftwf_plan_dft_c2r_2d fftwf_execute_dft_r2c fftwf_execute_dft_c2r
 *   fftwf_destroy_plan fftwf_cleanup
 * with the semantics of the FFTW 3 manual (row-major, r2c output n0 x (n1/2+1),
 * FFTW_FORWARD = -1, FFTW_BACKWARD = +1, unnormalised, new-array execute is
 * thread-safe).  Implementation: oracle/fftw_shim/fftw_shim.cpp on top of
 * oracle/fft_core.hpp.
 */
#ifndef ORACLE_FFTW3_SHIM_H
#define ORACLE_FFTW3_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef float fftwf_complex[2];
typedef struct fftwf_plan_s *fftwf_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_DESTROY_INPUT (1U << 0)
#define FFTW_ESTIMATE (1U << 6)

void *fftwf_malloc(size_t n);
void fftwf_free(void *p);
fftwf_plan fftwf_plan_dft_2d(int n0, int n1, fftwf_complex *in,
                             fftwf_complex *out, int sign, unsigned flags);
fftwf_plan fftwf_plan_dft_r2c_2d(int n0, int n1, float *in, fftwf_complex *out,
                                 unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_2d(int n0, int n1, fftwf_complex *in, float *out,
                                 unsigned flags);
void fftwf_execute_dft_r2c(const fftwf_plan p, float *in, fftwf_complex *out);
void fftwf_execute_dft_c2r(const fftwf_plan p, fftwf_complex *in, float *out);
void fftwf_execute_dft(const fftwf_plan p, fftwf_complex *in, fftwf_complex *out);
void fftwf_destroy_plan(fftwf_plan p);
void fftwf_cleanup(void);

#ifdef __cplusplus
}
#endif
#endif
