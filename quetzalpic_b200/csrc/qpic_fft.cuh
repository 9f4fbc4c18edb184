// Hand-written batched shared-memory FFTs (FP64) for the transverse solves, fused with the spectral division.
//
// Transform kinds = what the reference asks of FFTW (src/Fourier2d.cpp:32-43): unnormalised 2-D r2c (forward, sign -1)
// and c2r over an (ny, nz) real array, half spectrum along z.  One solve (src/System_3d.cpp:686-706) is
//     rows forward  ->  [columns forward -> / symbol -> columns backward]  ->  rows backward, then / ny / nz.
// The bracket runs inside ONE kernel with the columns resident in shared memory, so the spectrum is read and
// written once instead of three times.
//
//   * psi solves (real field): two real rows ride one complex row transform (r = row 2m, i = row 2m+1), split into
//     the two Hermitian half spectra in shared memory, columns over nz/2+1 bins -- same arithmetic volume as r2c/c2r.
//   * By and Bz (src/System_3d.cpp:529-580) have the same operator (D = magnetic_field_D) and independent sources, so
//     they are solved as ONE complex field By + i*Bz: full-length row transforms, nz columns, no split at all.  The
//     sources (:530-536, :555-564) are assembled while loading the rows; the results, the /ny/nz scaling and the
//     |B| > threshold / NaN stop test (:541-552, :569-580) are the epilogue of the backward row kernel.
//
// 1-D engine: Stockham autosort, radix 4 / 2 specialised, any other prime factor through a generic O(r^2) butterfly
// (200 = 4*2*5*5 and 100 are the shipped grid sizes; 23 appears in the test fixtures).  Twiddles come from one table
// exp(-2 pi i t / n) per length, computed on the host in extended precision.
#pragma once

#include "qpic_compat.h"
#include "qpic_kernels.cuh"

namespace qpic {

constexpr int FFT_MAX_FACTORS = 20;

struct FftPlan {
    int n;
    int nfac;
    int fac[FFT_MAX_FACTORS];
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

template <bool INV>
__device__ __forceinline__ double2 twiddle(const double2 * __restrict__ tw, int t) {
    double2 w = tw[t];
    if (INV) w.y = -w.y;
    return w;
}

// `nb` independent length-n sequences stored back to back in a[] (sequence b at a + b*n); w[] is scratch of the same
// size.  Called by every thread of the block; returns the buffer that holds the result (the other one is free).
template <bool INV>
__device__ double2 * fft_batch_smem(double2 * a, double2 * w, int nb, const FftPlan & P, const double2 * __restrict__ tw) {
    const int n = P.n;
    int s = 1, ncur = n;
    for (int f = 0; f < P.nfac; f++) {
        const int r = P.fac[f];
        const int m = ncur / r;
        const int nbf = n / r;  // butterflies per sequence; element j of butterfly u sits at u + nbf*j
        const int total = nb * nbf;
        if (r == 4) {
            for (int t = threadIdx.x; t < total; t += blockDim.x) {
                const int b = t / nbf, u = t - b * nbf;
                const int p = u / s, q = u - p * s;
                const double2 * x = a + b * n;
                double2 * y = w + b * n + q + s * 4 * p;
                const double2 a0 = x[u], a1 = x[u + nbf], a2 = x[u + 2 * nbf], a3 = x[u + 3 * nbf];
                const double2 t0 = make_double2(a0.x + a2.x, a0.y + a2.y);
                const double2 t1 = make_double2(a0.x - a2.x, a0.y - a2.y);
                const double2 t2 = make_double2(a1.x + a3.x, a1.y + a3.y);
                const double dr = a1.x - a3.x, di = a1.y - a3.y;
                // (a1 - a3) * (-i) forward, * (+i) backward
                const double2 t3 = INV ? make_double2(-di, dr) : make_double2(di, -dr);
                y[0] = make_double2(t0.x + t2.x, t0.y + t2.y);
                if (p == 0) {
                    y[s] = make_double2(t1.x + t3.x, t1.y + t3.y);
                    y[2 * s] = make_double2(t0.x - t2.x, t0.y - t2.y);
                    y[3 * s] = make_double2(t1.x - t3.x, t1.y - t3.y);
                } else {
                    y[s] = cmul(make_double2(t1.x + t3.x, t1.y + t3.y), twiddle<INV>(tw, p * s));
                    y[2 * s] = cmul(make_double2(t0.x - t2.x, t0.y - t2.y), twiddle<INV>(tw, 2 * p * s));
                    y[3 * s] = cmul(make_double2(t1.x - t3.x, t1.y - t3.y), twiddle<INV>(tw, 3 * p * s));
                }
            }
        } else if (r == 2) {
            for (int t = threadIdx.x; t < total; t += blockDim.x) {
                const int b = t / nbf, u = t - b * nbf;
                const int p = u / s, q = u - p * s;
                const double2 * x = a + b * n;
                double2 * y = w + b * n + q + s * 2 * p;
                const double2 a0 = x[u], a1 = x[u + nbf];
                y[0] = make_double2(a0.x + a1.x, a0.y + a1.y);
                const double2 d = make_double2(a0.x - a1.x, a0.y - a1.y);
                y[s] = (p == 0) ? d : cmul(d, twiddle<INV>(tw, p * s));
            }
        } else {
            // generic radix r: y_k = w_p^k * sum_j x_j W_r^{jk},  W_r^t = tw[t * nbf]
            for (int t = threadIdx.x; t < total; t += blockDim.x) {
                const int b = t / nbf, u = t - b * nbf;
                const int p = u / s, q = u - p * s;
                const double2 * x = a + b * n;
                double2 * y = w + b * n + q + s * r * p;
                for (int k = 0; k < r; k++) {
                    double2 acc = x[u];
                    int jk = 0;
                    for (int j = 1; j < r; j++) {
                        jk += k;
                        if (jk >= r) jk -= r;
                        const double2 c = twiddle<INV>(tw, jk * nbf);
                        const double2 xv = x[u + nbf * j];
                        acc.x += xv.x * c.x - xv.y * c.y;
                        acc.y += xv.x * c.y + xv.y * c.x;
                    }
                    y[s * k] = (p == 0 || k == 0) ? acc : cmul(acc, twiddle<INV>(tw, p * k * s));
                }
            }
        }
        __syncthreads();
        double2 * tmp = a;
        a = w;
        w = tmp;
        s *= r;
        ncur = m;
    }
    return a;
}

// ------------------------------------------------------------------------------------------------ real field (psi)

// forward rows, two real rows per block -> half spectra spec[ny][nzh]
__global__ void k_fft_rows_r2c(const Ctl * ctl, FftPlan Pz, const double2 * __restrict__ twz, const double * __restrict__ src, double2 * __restrict__ spec,
                               int ny, int nz) {
    if (ctl->stop_tag) return;
    QPIC_DYN_SMEM(double2, sm);
    double2 * a = sm;
    double2 * w = sm + nz;
    const int nzh = nz / 2 + 1;
    const int j0 = 2 * blockIdx.x, j1 = j0 + 1;
    const bool has1 = j1 < ny;
    for (int k = threadIdx.x; k < nz; k += blockDim.x) {
        a[k] = make_double2(src[(size_t)j0 * nz + k], has1 ? src[(size_t)j1 * nz + k] : 0.0);
    }
    __syncthreads();
    const double2 * z = fft_batch_smem<false>(a, w, 1, Pz, twz);
    for (int k = threadIdx.x; k < nzh; k += blockDim.x) {
        const double2 zk = z[k];
        const double2 zm = z[k == 0 ? 0 : nz - k];
        // U0 = (Z_k + conj(Z_{n-k})) / 2,  U1 = (Z_k - conj(Z_{n-k})) / (2i)
        spec[(size_t)j0 * nzh + k] = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
        if (has1) spec[(size_t)j1 * nzh + k] = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));
    }
}

// backward rows: two half spectra -> one complex sequence -> two real rows, scaled by 1/ny/nz (src/System_3d.cpp:361)
__global__ void k_fft_rows_c2r(const Ctl * ctl, FftPlan Pz, const double2 * __restrict__ twz, const double2 * __restrict__ spec, double * __restrict__ dst,
                               int ny, int nz) {
    if (ctl->stop_tag) return;
    QPIC_DYN_SMEM(double2, sm);
    double2 * a = sm;
    double2 * w = sm + nz;
    const int nzh = nz / 2 + 1;
    const int j0 = 2 * blockIdx.x, j1 = j0 + 1;
    const bool has1 = j1 < ny;
    for (int k = threadIdx.x; k < nzh; k += blockDim.x) {
        double2 v0 = spec[(size_t)j0 * nzh + k];
        double2 v1 = has1 ? spec[(size_t)j1 * nzh + k] : make_double2(0.0, 0.0);
        const int km = nz - k;
        if (k == 0 || km == k) {
            // c2r ignores the imaginary part of the DC / Nyquist bins (they are rounding noise for a real field)
            v0.y = 0.0;
            v1.y = 0.0;
        }
        a[k] = make_double2(v0.x - v1.y, v0.y + v1.x);  // V0 + i V1
        if (k != 0 && km != k) a[km] = make_double2(v0.x + v1.y, v1.x - v0.y);  // conj(V0) + i conj(V1)
    }
    __syncthreads();
    const double2 * z = fft_batch_smem<true>(a, w, 1, Pz, twz);
    for (int k = threadIdx.x; k < nz; k += blockDim.x) {
        const double2 v = z[k];
        dst[(size_t)j0 * nz + k] = v.x / ny / nz;
        if (has1) dst[(size_t)j1 * nz + k] = v.y / ny / nz;
    }
}

// ------------------------------------------------------------------------------------------------ columns + symbol
// spec has `ncols` columns with leading dimension ld; `cpb` adjacent columns per block.  The divisor is
// -D + (2cos(2 pi j/ny)-2)/dy/dy + (2cos(2 pi k/nz)-2)/dz/dz (src/System_3d.cpp:699), from two host-built tables
// sy[j], sz[k] (same libm cos as the reference); bin (0,0) is forced to zero (:688-689, :695-697).  For a full
// (non-half) spectrum the z table is indexed by min(k, nz-k), i.e. the Hermitian partner the reference divides.
__global__ void k_fft_cols_solve(const Ctl * ctl, FftPlan Py, const double2 * __restrict__ twy, double2 * __restrict__ spec, int ld, int ncols, int cpb,
                                 const double * __restrict__ sy, const double * __restrict__ sz, int nz_full, double D) {
    if (ctl->stop_tag) return;
    QPIC_DYN_SMEM(double2, sm);
    const int ny = Py.n;
    double2 * a = sm;
    double2 * w = sm + cpb * ny;
    const int k0 = blockIdx.x * cpb;
    const int nb = (ncols - k0 < cpb) ? ncols - k0 : cpb;
    const int total = nb * ny;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int j = t / nb, c = t - j * nb;
        a[c * ny + j] = spec[(size_t)j * ld + k0 + c];
    }
    __syncthreads();
    double2 * f = fft_batch_smem<false>(a, w, nb, Py, twy);
    double2 * other = (f == a) ? w : a;
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int c = t / ny, j = t - c * ny;
        const int k = k0 + c;
        const int kk = (k <= nz_full - k) ? k : nz_full - k;
        double2 v = f[t];
        if (j == 0 && k == 0) {
            v = make_double2(0.0, 0.0);
        } else {
            const double multiplier = -D + sy[j] + sz[kk];
            v.x /= multiplier;
            v.y /= multiplier;
        }
        f[t] = v;
    }
    __syncthreads();
    const double2 * res = fft_batch_smem<true>(f, other, nb, Py, twy);
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int j = t / nb, c = t - j * nb;
        spec[(size_t)j * ld + k0 + c] = res[c * ny + j];
    }
}

// ------------------------------------------------------------------------------------------------ complex field (By + i Bz)

// forward rows of the B solve; the right-hand sides are built on load:
//   re: -D*by(i) - (jx(j,k+1)-jx(j,k))/dz + djz_dxi,  djz_dxi = (jz - jz_next)/dx    (src/System_3d.cpp:530-536, :489-495)
//   im: -D*bz(i) + (jx(j+1,k)-jx(j,k))/dy - djy_dxi,  djy_dxi = (jy - jy_next)/dx    (:555-564)
// NB the last row takes bz(i, 0, k) instead of bz(i, ny-1, k) -- src/System_3d.cpp:563, reproduced on purpose.
__global__ void k_fft_rows_b_forward(const Ctl * ctl, FftPlan Pz, const double2 * __restrict__ twz, Geom g, double D, const double * __restrict__ by,
                                     const double * __restrict__ bz, const double * __restrict__ jx, const double * __restrict__ jy,
                                     const double * __restrict__ jz, const double * __restrict__ jy_next, const double * __restrict__ jz_next,
                                     double2 * __restrict__ spec) {
    if (ctl->stop_tag) return;
    QPIC_DYN_SMEM(double2, sm);
    const int nz = g.nz, ny = g.ny;
    double2 * a = sm;
    double2 * w = sm + nz;
    const int j = blockIdx.x;
    const int jp = wrap_inc(j, ny);
    const int jq = (j == ny - 1) ? 0 : j;
    for (int k = threadIdx.x; k < nz; k += blockDim.x) {
        const int t = j * nz + k;
        const int kp = wrap_inc(k, nz);
        const double djz_dxi = (jz[t] - jz_next[t]) / g.dx;
        const double djy_dxi = (jy[t] - jy_next[t]) / g.dx;
        const double sre = -D * by[t] - (jx[j * nz + kp] - jx[t]) / g.dz + djz_dxi;
        const double sim = -D * bz[jq * nz + k] + (jx[jp * nz + k] - jx[t]) / g.dy - djy_dxi;
        a[k] = make_double2(sre, sim);
    }
    __syncthreads();
    const double2 * z = fft_batch_smem<false>(a, w, 1, Pz, twz);
    for (int k = threadIdx.x; k < nz; k += blockDim.x) spec[(size_t)j * nz + k] = z[k];
}

// backward rows of the B solve: by(i) = re/ny/nz, bz(i) = im/ny/nz and the stop test (src/System_3d.cpp:541-552, :569-580)
__global__ void k_fft_rows_b_backward(Ctl * ctl, FftPlan Pz, const double2 * __restrict__ twz, Geom g, const double2 * __restrict__ spec,
                                      double * __restrict__ by, double * __restrict__ bz, double threshold, int iteration) {
    // blocks of THIS launch may already have raised the flag; only an earlier launch's flag cancels the work
    const int my_tag = ctl->slice * STOP_TAG_STRIDE + iteration + 1;
    const int tag = ctl->stop_tag;
    if (tag != 0 && tag != my_tag) return;
    QPIC_DYN_SMEM(double2, sm);
    const int nz = g.nz, ny = g.ny;
    double2 * a = sm;
    double2 * w = sm + nz;
    const int j = blockIdx.x;
    for (int k = threadIdx.x; k < nz; k += blockDim.x) a[k] = spec[(size_t)j * nz + k];
    __syncthreads();
    const double2 * z = fft_batch_smem<true>(a, w, 1, Pz, twz);
    bool bad = false;
    for (int k = threadIdx.x; k < nz; k += blockDim.x) {
        const double vy = z[k].x / ny / nz;
        const double vz = z[k].y / ny / nz;
        by[j * nz + k] = vy;
        bz[j * nz + k] = vz;
        // !(|v| <= threshold) is true for |v| > threshold and for NaN
        if (!(fabs(vy) <= threshold) || !(fabs(vz) <= threshold)) bad = true;
    }
    if (bad) atomicExch(&ctl->stop_tag, my_tag);
}

// generic complex rows (used by the exposed Helmholtz-pair building block and the tests)
__global__ void k_fft_rows_c2c(FftPlan Pz, const double2 * __restrict__ twz, const double * __restrict__ re, const double * __restrict__ im,
                               double2 * __restrict__ spec, int nz) {
    QPIC_DYN_SMEM(double2, sm);
    double2 * a = sm;
    double2 * w = sm + nz;
    const int j = blockIdx.x;
    for (int k = threadIdx.x; k < nz; k += blockDim.x) a[k] = make_double2(re[(size_t)j * nz + k], im[(size_t)j * nz + k]);
    __syncthreads();
    const double2 * z = fft_batch_smem<false>(a, w, 1, Pz, twz);
    for (int k = threadIdx.x; k < nz; k += blockDim.x) spec[(size_t)j * nz + k] = z[k];
}

}  // namespace qpic
