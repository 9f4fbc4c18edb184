/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * Plain-C mixed-radix FFT used (a) by the CPU oracle restatement (oracle/qpic_oracle.c) and (b) as the
 * backend of the `fftw3.h` shim that lets the unmodified reference compile (oracle/shims/fftw3.h).
 *
 * The reference's only third-party arithmetic is FFTW3 (version unpinned, CMakeLists.txt:30), called
 * through src/Fourier2d.cpp:32-43 as fftw_plan_dft_r2c_2d / fftw_plan_dft_c2r_2d: unnormalised 2-D
 * real<->half-complex DFTs, forward sign -1, half spectrum along the LAST axis.  That transform is
 * mathematically defined, so this file restates it from the definition (Stockham autosort, radices
 * 4/2/3/5/7 + a generic odd-prime butterfly) and tests/test_oracle_fft.py pins it against numpy.fft.
 */
#ifndef QO_FFT_H
#define QO_FFT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int n;
    int nfac;
    int fac[32];
    double * tw; /* n complex values exp(-2 pi i t / n), interleaved re,im */
} qo_fft1d;

/* returns 0 on success */
int qo_fft1d_init(qo_fft1d * plan, int n);
void qo_fft1d_free(qo_fft1d * plan);
/* unnormalised complex DFT of n interleaved values, in place; work holds n complex values.
 * sign = -1: forward (exp(-i...)), sign = +1: backward. */
void qo_fft1d_exec(const qo_fft1d * plan, double * x, double * work, int sign);

typedef struct {
    int n1, n2;
    qo_fft1d p1, p2;
} qo_fft2d;

int qo_fft2d_init(qo_fft2d * plan, int n1, int n2);
void qo_fft2d_free(qo_fft2d * plan);
/* in: n1*n2 reals (row-major, n2 fastest).  out: n1*(n2/2+1) complex, interleaved. */
void qo_rfft2_forward(const qo_fft2d * plan, const double * in, double * out);
/* inverse of the above, unnormalised (result = n1*n2 * original).  `out` is used as scratch. */
void qo_rfft2_backward(const qo_fft2d * plan, double * out, double * in);

#ifdef __cplusplus
}
#endif

#endif
