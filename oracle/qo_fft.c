/* TEST INFRASTRUCTURE ONLY -- see qo_fft.h.  Stockham autosort mixed-radix complex FFT and the
 * 2-D real<->half-complex transforms with FFTW's r2c/c2r conventions (reference call sites:
 * src/Fourier2d.cpp:32-43). */
#include "qo_fft.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int qo_fft1d_init(qo_fft1d * plan, int n) {
    if (n < 1) return -1;
    plan->n = n;
    plan->nfac = 0;
    int m = n;
    while (m % 4 == 0) {
        plan->fac[plan->nfac++] = 4;
        m /= 4;
    }
    while (m % 2 == 0) {
        plan->fac[plan->nfac++] = 2;
        m /= 2;
    }
    for (int f = 3; m > 1; f += 2) {
        while (m % f == 0) {
            if (plan->nfac >= 32) return -1;
            plan->fac[plan->nfac++] = f;
            m /= f;
        }
    }
    plan->tw = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    if (!plan->tw) return -1;
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int t = 0; t < n; t++) {
        const long double a = two_pi * (long double)t / (long double)n;
        plan->tw[2 * t] = (double)cosl(a);
        plan->tw[2 * t + 1] = (double)(-sinl(a));
    }
    return 0;
}

void qo_fft1d_free(qo_fft1d * plan) {
    free(plan->tw);
    plan->tw = NULL;
}

/* one Stockham pass of radix r:  out[q + s*(r*p + k)] = w_p^k * sum_j in[q + s*(p + m*j)] * W_r^{jk}  */
static void stockham_pass(const double * in, double * out, int n, int s, int m, int r, const double * tw, int sign) {
    const double sg = (sign < 0) ? 1.0 : -1.0; /* multiplies the stored (forward) twiddle's imaginary part */
    if (r == 2) {
        for (int p = 0; p < m; p++) {
            const double wr = tw[2 * (p * s)], wi = sg * tw[2 * (p * s) + 1];
            for (int q = 0; q < s; q++) {
                const double * a0 = in + 2 * (q + s * p);
                const double * a1 = in + 2 * (q + s * (p + m));
                double * y0 = out + 2 * (q + s * (2 * p));
                double * y1 = out + 2 * (q + s * (2 * p + 1));
                const double dr = a0[0] - a1[0], di = a0[1] - a1[1];
                y0[0] = a0[0] + a1[0];
                y0[1] = a0[1] + a1[1];
                y1[0] = dr * wr - di * wi;
                y1[1] = dr * wi + di * wr;
            }
        }
        return;
    }
    if (r == 4) {
        for (int p = 0; p < m; p++) {
            const double w1r = tw[2 * (p * s)], w1i = sg * tw[2 * (p * s) + 1];
            const double w2r = tw[2 * (2 * p * s)], w2i = sg * tw[2 * (2 * p * s) + 1];
            const double w3r = tw[2 * (3 * p * s)], w3i = sg * tw[2 * (3 * p * s) + 1];
            for (int q = 0; q < s; q++) {
                const double * a0 = in + 2 * (q + s * p);
                const double * a1 = in + 2 * (q + s * (p + m));
                const double * a2 = in + 2 * (q + s * (p + 2 * m));
                const double * a3 = in + 2 * (q + s * (p + 3 * m));
                const double t0r = a0[0] + a2[0], t0i = a0[1] + a2[1];
                const double t1r = a0[0] - a2[0], t1i = a0[1] - a2[1];
                const double t2r = a1[0] + a3[0], t2i = a1[1] + a3[1];
                /* (a1 - a3) * (-i) forward, * (+i) backward */
                const double dr = a1[0] - a3[0], di = a1[1] - a3[1];
                const double t3r = sg * di, t3i = -sg * dr;
                double * y0 = out + 2 * (q + s * (4 * p));
                double * y1 = out + 2 * (q + s * (4 * p + 1));
                double * y2 = out + 2 * (q + s * (4 * p + 2));
                double * y3 = out + 2 * (q + s * (4 * p + 3));
                y0[0] = t0r + t2r;
                y0[1] = t0i + t2i;
                double br = t1r + t3r, bi = t1i + t3i;
                y1[0] = br * w1r - bi * w1i;
                y1[1] = br * w1i + bi * w1r;
                br = t0r - t2r;
                bi = t0i - t2i;
                y2[0] = br * w2r - bi * w2i;
                y2[1] = br * w2i + bi * w2r;
                br = t1r - t3r;
                bi = t1i - t3i;
                y3[0] = br * w3r - bi * w3i;
                y3[1] = br * w3i + bi * w3r;
            }
        }
        return;
    }
    /* generic radix: direct r-point DFT with table twiddles W_r^t = tw[t * (n / r)] */
    const int step = n / r;
    double ar[64], ai[64];
    if (r > 64) {
        /* handled by the caller: qo_fft1d_exec falls back for huge prime factors */
        return;
    }
    for (int p = 0; p < m; p++) {
        for (int q = 0; q < s; q++) {
            for (int j = 0; j < r; j++) {
                ar[j] = in[2 * (q + s * (p + m * j))];
                ai[j] = in[2 * (q + s * (p + m * j)) + 1];
            }
            for (int k = 0; k < r; k++) {
                double sr = ar[0], si = ai[0];
                for (int j = 1; j < r; j++) {
                    const int t = ((j * k) % r) * step;
                    const double cr = tw[2 * t], ci = sg * tw[2 * t + 1];
                    sr += ar[j] * cr - ai[j] * ci;
                    si += ar[j] * ci + ai[j] * cr;
                }
                const int t = p * k * s;
                const double wr = tw[2 * t], wi = sg * tw[2 * t + 1];
                out[2 * (q + s * (r * p + k))] = sr * wr - si * wi;
                out[2 * (q + s * (r * p + k)) + 1] = sr * wi + si * wr;
            }
        }
    }
}

/* O(n^2) definition, used only when a prime factor exceeds the generic butterfly's register array */
static void naive_dft(const double * in, double * out, int n, const double * tw, int sign) {
    const double sg = (sign < 0) ? 1.0 : -1.0;
    for (int k = 0; k < n; k++) {
        double sr = 0, si = 0;
        for (int j = 0; j < n; j++) {
            const int t = (int)(((long long)j * k) % n);
            const double cr = tw[2 * t], ci = sg * tw[2 * t + 1];
            sr += in[2 * j] * cr - in[2 * j + 1] * ci;
            si += in[2 * j] * ci + in[2 * j + 1] * cr;
        }
        out[2 * k] = sr;
        out[2 * k + 1] = si;
    }
}

void qo_fft1d_exec(const qo_fft1d * plan, double * x, double * work, int sign) {
    const int n = plan->n;
    for (int f = 0; f < plan->nfac; f++) {
        if (plan->fac[f] > 64) {
            naive_dft(x, work, n, plan->tw, sign);
            memcpy(x, work, sizeof(double) * 2 * (size_t)n);
            return;
        }
    }
    double * a = x;
    double * b = work;
    int s = 1, ncur = n;
    for (int f = 0; f < plan->nfac; f++) {
        const int r = plan->fac[f];
        const int m = ncur / r;
        stockham_pass(a, b, n, s, m, r, plan->tw, sign);
        double * t = a;
        a = b;
        b = t;
        s *= r;
        ncur = m;
    }
    if (a != x) memcpy(x, a, sizeof(double) * 2 * (size_t)n);
}

int qo_fft2d_init(qo_fft2d * plan, int n1, int n2) {
    plan->n1 = n1;
    plan->n2 = n2;
    if (qo_fft1d_init(&plan->p1, n1) != 0) return -1;
    if (qo_fft1d_init(&plan->p2, n2) != 0) return -1;
    return 0;
}

void qo_fft2d_free(qo_fft2d * plan) {
    qo_fft1d_free(&plan->p1);
    qo_fft1d_free(&plan->p2);
}

void qo_rfft2_forward(const qo_fft2d * plan, const double * in, double * out) {
    const int n1 = plan->n1, n2 = plan->n2, nh = n2 / 2 + 1;
    const int nmax = n1 > n2 ? n1 : n2;
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        double * buf = (double *)malloc(sizeof(double) * 4 * (size_t)nmax);
        double * work = buf + 2 * (size_t)nmax;
#ifdef _OPENMP
#pragma omp for
#endif
        for (int i = 0; i < n1; i++) {
            for (int k = 0; k < n2; k++) {
                buf[2 * k] = in[(size_t)n2 * i + k];
                buf[2 * k + 1] = 0.0;
            }
            qo_fft1d_exec(&plan->p2, buf, work, -1);
            memcpy(out + 2 * (size_t)nh * i, buf, sizeof(double) * 2 * (size_t)nh);
        }
#ifdef _OPENMP
#pragma omp for
#endif
        for (int k = 0; k < nh; k++) {
            for (int i = 0; i < n1; i++) {
                buf[2 * i] = out[2 * ((size_t)nh * i + k)];
                buf[2 * i + 1] = out[2 * ((size_t)nh * i + k) + 1];
            }
            qo_fft1d_exec(&plan->p1, buf, work, -1);
            for (int i = 0; i < n1; i++) {
                out[2 * ((size_t)nh * i + k)] = buf[2 * i];
                out[2 * ((size_t)nh * i + k) + 1] = buf[2 * i + 1];
            }
        }
        free(buf);
    }
}

void qo_rfft2_backward(const qo_fft2d * plan, double * out, double * in) {
    const int n1 = plan->n1, n2 = plan->n2, nh = n2 / 2 + 1;
    const int nmax = n1 > n2 ? n1 : n2;
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        double * buf = (double *)malloc(sizeof(double) * 4 * (size_t)nmax);
        double * work = buf + 2 * (size_t)nmax;
#ifdef _OPENMP
#pragma omp for
#endif
        for (int k = 0; k < nh; k++) {
            for (int i = 0; i < n1; i++) {
                buf[2 * i] = out[2 * ((size_t)nh * i + k)];
                buf[2 * i + 1] = out[2 * ((size_t)nh * i + k) + 1];
            }
            qo_fft1d_exec(&plan->p1, buf, work, +1);
            for (int i = 0; i < n1; i++) {
                out[2 * ((size_t)nh * i + k)] = buf[2 * i];
                out[2 * ((size_t)nh * i + k) + 1] = buf[2 * i + 1];
            }
        }
#ifdef _OPENMP
#pragma omp for
#endif
        for (int i = 0; i < n1; i++) {
            /* Hermitian extension along the last axis, then a complex inverse; the real part is the
             * c2r result (imaginary parts of the DC / Nyquist bins drop out, as in FFTW). */
            const double * row = out + 2 * (size_t)nh * i;
            for (int k = 0; k < nh; k++) {
                buf[2 * k] = row[2 * k];
                buf[2 * k + 1] = row[2 * k + 1];
            }
            for (int k = nh; k < n2; k++) {
                buf[2 * k] = row[2 * (n2 - k)];
                buf[2 * k + 1] = -row[2 * (n2 - k) + 1];
            }
            qo_fft1d_exec(&plan->p2, buf, work, +1);
            for (int k = 0; k < n2; k++) in[(size_t)n2 * i + k] = buf[2 * k];
        }
        free(buf);
    }
}
