/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * Stand-in for <fftw3.h> (FFTW is not installed in this image).  It declares exactly the FFTW entry
 * points the UNMODIFIED reference calls (src/Fourier2d.cpp:4-5,10-11,14,17,33-34,38,42;
 * src/quasistatic_pic.cpp:29-30; src/Fourier2d.h:2,7,16-17); fftw3_shim.cpp implements them on top of
 * oracle/qo_fft.c (plain mixed-radix FFT, pinned against numpy.fft) or, when QPIC_SHIM_MKL is defined at
 * build time, on top of the MKL DFTI symbols that torch's libtorch_cpu.so exports (timed CPU baseline).
 */
#ifndef QPIC_SHIM_FFTW3_H
#define QPIC_SHIM_FFTW3_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef struct qpic_shim_plan_s * fftw_plan;

#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

double * fftw_alloc_real(size_t n);
fftw_complex * fftw_alloc_complex(size_t n);
void fftw_free(void * p);

fftw_plan fftw_plan_dft_r2c_2d(int n0, int n1, double * in, fftw_complex * out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_2d(int n0, int n1, fftw_complex * in, double * out, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);

int fftw_init_threads(void);
void fftw_plan_with_nthreads(int nthreads);

#ifdef __cplusplus
}
#endif

#endif
