// Hand-written batched shared-memory FFTs (FP64) for the transverse solves, fused with the spectral division.
//
// Transform kinds = what the reference asks of FFTW (src/Fourier2d.cpp:32-43): unnormalised 2-D r2c (forward, sign -1)
// and c2r over an (ny, nz) real array, half spectrum along z.  One solve (src/System_3d.cpp:686-706) is
//     rows forward  ->  [columns forward -> / symbol -> columns backward]  ->  rows backward, then / ny / nz.
// The bracket runs inside ONE kernel with the columns resident in shared memory, so the spectrum is read and
// written once instead of three times.
//
//   * psi solves (real field): two real rows ride one complex row transform (re = row 2m, im = row 2m+1), split into
//     the two Hermitian half spectra in shared memory, columns over nz/2+1 bins -- same arithmetic volume as r2c/c2r.
//   * By and Bz (src/System_3d.cpp:529-580) have the same operator (D = magnetic_field_D) and independent sources, so
//     they are solved as ONE complex field By + i*Bz: full-length row transforms, nz columns, no split at all.  The
//     sources (:530-536, :555-564) are assembled while loading the rows; the results, the /ny/nz scaling and the
//     |B| > threshold / NaN stop test (:541-552, :569-580) are the epilogue of the backward row kernel.
//
// 1-D engines (Stockham autosort, decimation in frequency: reads are unit stride, so bank-conflict free):
//   * power-of-two lengths >= 64 (512, 1024, 4096: the BASELINE grids): every thread keeps 8 points in registers and
//     does radix-8 passes (plus one radix-4/2 pass to finish), in place in shared memory, index math by shifts;
//   * any other length (200 = 4*2*5*5 and 100 are the shipped grids, 23 appears in the fixtures): radix 4 / 2 and a
//     generic O(r^2) butterfly for other prime factors, ping-pong buffers.
// Twiddles come from one table exp(-2 pi i t / n) per length, computed on the host in extended precision.
#pragma once

#include "qpic_compat.h"
#include "qpic_kernels.cuh"

namespace qpic {

constexpr int FFT_MAX_FACTORS = 20;

struct FftPlan {
    int n;
    int nfac;
    int fac[FFT_MAX_FACTORS];
    int log2n;  // > 0: power of two handled by the register engine; 0: generic engine
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiplication by -i (forward) / +i (backward)
template <bool INV>
__device__ __forceinline__ double2 rot90(double2 a) {
    return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

template <bool INV>
__device__ __forceinline__ double2 twiddle(const double2 * __restrict__ tw, int t) {
    double2 w = tw[t];
    if (INV) w.y = -w.y;
    return w;
}

// ------------------------------------------------------------------------------------------------ generic engine

// `nb` independent length-n sequences stored back to back in a[] (sequence b at a + b*n); w[] is scratch of the same
// size.  Called by every thread of the block; returns the buffer that holds the result (the other one is free).
template <bool INV>
__device__ double2 * fft_generic_smem(double2 * a, double2 * w, int nb, const FftPlan & P, const double2 * __restrict__ tw) {
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
                const double2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3);
                const double2 t3 = rot90<INV>(csub(a1, a3));
                y[0] = cadd(t0, t2);
                if (p == 0) {
                    y[s] = cadd(t1, t3);
                    y[2 * s] = csub(t0, t2);
                    y[3 * s] = csub(t1, t3);
                } else {
                    y[s] = cmul(cadd(t1, t3), twiddle<INV>(tw, p * s));
                    y[2 * s] = cmul(csub(t0, t2), twiddle<INV>(tw, 2 * p * s));
                    y[3 * s] = cmul(csub(t1, t3), twiddle<INV>(tw, 3 * p * s));
                }
            }
        } else if (r == 2) {
            for (int t = threadIdx.x; t < total; t += blockDim.x) {
                const int b = t / nbf, u = t - b * nbf;
                const int p = u / s, q = u - p * s;
                const double2 * x = a + b * n;
                double2 * y = w + b * n + q + s * 2 * p;
                const double2 a0 = x[u], a1 = x[u + nbf];
                y[0] = cadd(a0, a1);
                const double2 d = csub(a0, a1);
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

// ------------------------------------------------------------------------------------------------ power-of-two engine

template <bool INV>
__device__ __forceinline__ void dft4(double2 & x0, double2 & x1, double2 & x2, double2 & x3) {
    const double2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3);
    const double2 t3 = rot90<INV>(csub(x1, x3));
    x0 = cadd(t0, t2);
    x1 = cadd(t1, t3);
    x2 = csub(t0, t2);
    x3 = csub(t1, t3);
}

// 8-point DFT, natural order in and out
template <bool INV>
__device__ __forceinline__ void dft8(double2 (&v)[8]) {
    constexpr double H = 0.70710678118654752440;  // sqrt(1/2)
    const double2 b0 = cadd(v[0], v[4]), b1 = cadd(v[1], v[5]), b2 = cadd(v[2], v[6]), b3 = cadd(v[3], v[7]);
    double2 c0 = csub(v[0], v[4]), c1 = csub(v[1], v[5]), c2 = csub(v[2], v[6]), c3 = csub(v[3], v[7]);
    // c_j *= W8^j:  W8 = H(1 - i) forward, H(1 + i) backward; W8^2 = -/+ i; W8^3 = H(-1 - i) / H(-1 + i)
    c1 = INV ? make_double2(H * (c1.x - c1.y), H * (c1.x + c1.y)) : make_double2(H * (c1.x + c1.y), H * (c1.y - c1.x));
    c2 = rot90<INV>(c2);
    c3 = INV ? make_double2(-H * (c3.x + c3.y), H * (c3.x - c3.y)) : make_double2(H * (c3.y - c3.x), -H * (c3.x + c3.y));
    double2 e0 = b0, e1 = b1, e2 = b2, e3 = b3;
    dft4<INV>(e0, e1, e2, e3);  // even outputs Y0 Y2 Y4 Y6
    dft4<INV>(c0, c1, c2, c3);  // odd outputs  Y1 Y3 Y5 Y7
    v[0] = e0;
    v[1] = c0;
    v[2] = e1;
    v[3] = c1;
    v[4] = e2;
    v[5] = c2;
    v[6] = e3;
    v[7] = c3;
}

// In-place transform of the sequences stored back to back in x[] (length n = 2^log2n >= 64 each).  The block must have
// exactly nb * n/8 threads: thread (b, t) owns points t + (n/8) e, e = 0..7, of sequence b in every pass.  The caller
// must __syncthreads() between filling x[] and this call; the result is visible to the whole block on return.
template <bool INV>
__device__ __forceinline__ void fft_pow2_smem(double2 * x, int log2n, const double2 * __restrict__ tw) {
    const int le8 = log2n - 3;
    const int e8 = 1 << le8;
    double2 * xs = x + ((threadIdx.x >> le8) << log2n);
    const int t = threadIdx.x & (e8 - 1);
    int log2s = 0;
    for (int rem = log2n; rem > 0;) {
        const int lr = rem >= 3 ? 3 : rem;  // log2(radix) of this pass
        double2 v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = xs[t + (e << le8)];
        if (lr == 3) {
            dft8<INV>(v);
        } else if (lr == 2) {
            dft4<INV>(v[0], v[2], v[4], v[6]);
            dft4<INV>(v[1], v[3], v[5], v[7]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double2 a0 = v[i], a1 = v[i + 4];
                v[i] = cadd(a0, a1);
                v[i + 4] = csub(a0, a1);
            }
        }
        __syncthreads();  // every thread has read its inputs: the buffer may be overwritten
        const int smask = (1 << log2s) - 1;
        if (lr == 3) {
            const int p = t >> log2s, q = t & smask;
            double2 * y = xs + q + ((8 * p) << log2s);
            y[0] = v[0];
            if (p == 0) {
#pragma unroll
                for (int k = 1; k < 8; k++) y[k << log2s] = v[k];
            } else {
#pragma unroll
                for (int k = 1; k < 8; k++) y[k << log2s] = cmul(v[k], twiddle<INV>(tw, (p * k) << log2s));
            }
        } else if (lr == 2) {
            // the finishing pass always has m = 1, i.e. p = 0: no twiddles
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int u = t + (i << le8);
                const int p = u >> log2s, q = u & smask;
                double2 * y = xs + q + ((4 * p) << log2s);
#pragma unroll
                for (int k = 0; k < 4; k++) y[k << log2s] = v[i + 2 * k];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int u = t + (i << le8);
                const int p = u >> log2s, q = u & smask;
                double2 * y = xs + q + ((2 * p) << log2s);
                y[0] = v[i];
                y[1 << log2s] = v[i + 4];
            }
        }
        __syncthreads();
        log2s += lr;
        rem -= lr;
    }
}

// dispatch: `a` holds nb sequences; POW2 works in place (w unused), the generic engine ping-pongs between a and w
template <bool INV, bool POW2>
__device__ __forceinline__ double2 * fft_smem(double2 * a, double2 * w, int nb, const FftPlan & P, const double2 * __restrict__ tw) {
    if (POW2) {
        fft_pow2_smem<INV>(a, P.log2n, tw);
        return a;
    }
    return fft_generic_smem<INV>(a, w, nb, P, tw);
}

// ------------------------------------------------------------------------------------------------ real field (psi)

// forward rows: each sequence carries two real rows -> half spectra spec[ny][nzh]; `spb` sequences per block
template <bool POW2>
__global__ void k_fft_rows_r2c(const Ctl * ctl, FftPlan Pz, const double2 * __restrict__ twz, const double * __restrict__ src, double2 * __restrict__ spec,
                               int ny, int nz, int spb) {
    QPIC_KERNEL_BEGIN(KID_k_fft_rows_r2c);
    if (cancelled(ctl)) return;
    QPIC_DYN_SMEM(double2, sm);
    double2 * a = sm;
    double2 * w = sm + spb * nz;
    const int nzh = nz / 2 + 1;
    const int pair0 = blockIdx.x * spb;
    for (int b = 0; b < spb; b++) {
        const int j0 = 2 * (pair0 + b), j1 = j0 + 1;
        const double * r0 = src + (size_t)j0 * nz;
        const double * r1 = src + (size_t)j1 * nz;
        for (int k = threadIdx.x; k < nz; k += blockDim.x) {
            a[b * nz + k] = make_double2(j0 < ny ? r0[k] : 0.0, j1 < ny ? r1[k] : 0.0);
        }
    }
    __syncthreads();
    const double2 * zz = fft_smem<false, POW2>(a, w, spb, Pz, twz);
    for (int b = 0; b < spb; b++) {
        const int j0 = 2 * (pair0 + b), j1 = j0 + 1;
        if (j0 >= ny) continue;
        const double2 * z = zz + b * nz;
        for (int k = threadIdx.x; k < nzh; k += blockDim.x) {
            const double2 zk = z[k];
            const double2 zm = z[k == 0 ? 0 : nz - k];
            // U0 = (Z_k + conj(Z_{n-k})) / 2,  U1 = (Z_k - conj(Z_{n-k})) / (2i)
            spec[(size_t)j0 * nzh + k] = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
            if (j1 < ny) spec[(size_t)j1 * nzh + k] = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));
        }
    }
}

// x / ny / nz as the reference writes it (src/System_3d.cpp:361); when ny*nz is a power of two the same value is
// obtained, bit for bit, by one multiplication with the exactly representable 1/(ny*nz) (inv_scale > 0)
__device__ __forceinline__ double scale_result(double v, int ny, int nz, double inv_scale) {
    return inv_scale > 0.0 ? v * inv_scale : v / ny / nz;
}

// backward rows: two half spectra -> one complex sequence -> two real rows, scaled by 1/ny/nz (src/System_3d.cpp:361)
template <bool POW2>
__global__ void k_fft_rows_c2r(const Ctl * ctl, FftPlan Pz, const double2 * __restrict__ twz, const double2 * __restrict__ spec, double * __restrict__ dst,
                               int ny, int nz, int spb, double inv_scale) {
    QPIC_KERNEL_BEGIN(KID_k_fft_rows_c2r);
    if (cancelled(ctl)) return;
    QPIC_DYN_SMEM(double2, sm);
    double2 * a = sm;
    double2 * w = sm + spb * nz;
    const int nzh = nz / 2 + 1;
    const int pair0 = blockIdx.x * spb;
    for (int b = 0; b < spb; b++) {
        const int j0 = 2 * (pair0 + b), j1 = j0 + 1;
        double2 * z = a + b * nz;
        for (int k = threadIdx.x; k < nzh; k += blockDim.x) {
            double2 v0 = (j0 < ny) ? spec[(size_t)j0 * nzh + k] : make_double2(0.0, 0.0);
            double2 v1 = (j1 < ny) ? spec[(size_t)j1 * nzh + k] : make_double2(0.0, 0.0);
            const int km = nz - k;
            if (k == 0 || km == k) {
                // c2r ignores the imaginary part of the DC / Nyquist bins (they are rounding noise for a real field)
                v0.y = 0.0;
                v1.y = 0.0;
            }
            z[k] = make_double2(v0.x - v1.y, v0.y + v1.x);                          // V0 + i V1
            if (k != 0 && km != k) z[km] = make_double2(v0.x + v1.y, v1.x - v0.y);  // conj(V0) + i conj(V1)
        }
    }
    __syncthreads();
    const double2 * zz = fft_smem<true, POW2>(a, w, spb, Pz, twz);
    for (int b = 0; b < spb; b++) {
        const int j0 = 2 * (pair0 + b), j1 = j0 + 1;
        for (int k = threadIdx.x; k < nz; k += blockDim.x) {
            const double2 v = zz[b * nz + k];
            if (j0 < ny) dst[(size_t)j0 * nz + k] = scale_result(v.x, ny, nz, inv_scale);
            if (j1 < ny) dst[(size_t)j1 * nz + k] = scale_result(v.y, ny, nz, inv_scale);
        }
    }
}

// ------------------------------------------------------------------------------------------------ columns + symbol
// spec has `ncols` columns with leading dimension ld; `cpb` adjacent columns per block.  The divisor is
// -D + (2cos(2 pi j/ny)-2)/dy/dy + (2cos(2 pi k/nz)-2)/dz/dz (src/System_3d.cpp:699), from two host-built tables
// sy[j], sz[k] (same libm cos as the reference); bin (0,0) is forced to zero (:688-689, :695-697).  For a full
// (non-half) spectrum the z table is indexed by min(k, nz-k), i.e. the Hermitian partner the reference divides.
template <bool POW2>
__global__ void k_fft_cols_solve(const Ctl * ctl, FftPlan Py, const double2 * __restrict__ twy, double2 * __restrict__ spec, int ld, int ncols, int cpb,
                                 const double * __restrict__ sy, const double * __restrict__ sz, int nz_full, double D) {
    QPIC_KERNEL_BEGIN(KID_k_fft_cols_solve);
    if (cancelled(ctl)) return;
    QPIC_DYN_SMEM(double2, sm);
    const int ny = Py.n;
    double2 * a = sm;
    double2 * w = sm + cpb * ny;
    const int k0 = blockIdx.x * cpb;
    const int nb = (ncols - k0 < cpb) ? ncols - k0 : cpb;
    if (cpb == 1) {
        for (int j = threadIdx.x; j < ny; j += blockDim.x) a[j] = spec[(size_t)j * ld + k0];
    } else if (cpb == 2) {
        for (int t = threadIdx.x; t < 2 * ny; t += blockDim.x) {
            const int j = t >> 1, c = t & 1;
            a[c * ny + j] = (c < nb) ? spec[(size_t)j * ld + k0 + c] : make_double2(0.0, 0.0);
        }
    } else {
        for (int t = threadIdx.x; t < cpb * ny; t += blockDim.x) {
            const int j = t / cpb, c = t - j * cpb;
            a[c * ny + j] = (c < nb) ? spec[(size_t)j * ld + k0 + c] : make_double2(0.0, 0.0);
        }
    }
    __syncthreads();
    double2 * f = fft_smem<false, POW2>(a, w, cpb, Py, twy);
    double2 * other = (f == a) ? w : a;
    for (int c = 0; c < cpb; c++) {
        const int k = k0 + c;
        const int kk = (k <= nz_full - k) ? k : nz_full - k;
        const double base = (c < nb) ? sz[kk] : 0.0;
        for (int j = threadIdx.x; j < ny; j += blockDim.x) {
            double2 v = f[c * ny + j];
            if ((j == 0 && k == 0) || c >= nb) {
                v = make_double2(0.0, 0.0);
            } else {
                const double multiplier = -D + sy[j] + base;
                v.x /= multiplier;
                v.y /= multiplier;
            }
            f[c * ny + j] = v;
        }
    }
    __syncthreads();
    const double2 * res = fft_smem<true, POW2>(f, other, cpb, Py, twy);
    if (cpb == 1) {
        for (int j = threadIdx.x; j < ny; j += blockDim.x) spec[(size_t)j * ld + k0] = res[j];
    } else if (cpb == 2) {
        for (int t = threadIdx.x; t < 2 * ny; t += blockDim.x) {
            const int j = t >> 1, c = t & 1;
            if (c < nb) spec[(size_t)j * ld + k0 + c] = res[c * ny + j];
        }
    } else {
        for (int t = threadIdx.x; t < cpb * ny; t += blockDim.x) {
            const int j = t / cpb, c = t - j * cpb;
            if (c < nb) spec[(size_t)j * ld + k0 + c] = res[c * ny + j];
        }
    }
}

// ------------------------------------------------------------------------------------------------ complex field (By + i Bz)

// forward rows of the B solve (`spb` rows per block); the right-hand sides are built on load:
//   re: -D*by(i) - (jx(j,k+1)-jx(j,k))/dz + djz_dxi,  djz_dxi = (jz - jz_next)/dx    (src/System_3d.cpp:530-536, :489-495)
//   im: -D*bz(i) + (jx(j+1,k)-jx(j,k))/dy - djy_dxi,  djy_dxi = (jy - jy_next)/dx    (:555-564)
// NB the last row takes bz(i, 0, k) instead of bz(i, ny-1, k) -- src/System_3d.cpp:563, reproduced on purpose.
template <bool POW2>
__global__ void k_fft_rows_b_forward(const Ctl * ctl, FftPlan Pz, const double2 * __restrict__ twz, Geom g, double D, const double * __restrict__ by,
                                     const double * __restrict__ bz, const double * __restrict__ jx, const double * __restrict__ jy,
                                     const double * __restrict__ jz, const double * __restrict__ jy_next, const double * __restrict__ jz_next,
                                     double2 * __restrict__ spec, int spb) {
    QPIC_KERNEL_BEGIN(KID_k_fft_rows_b_forward);
    if (cancelled(ctl)) return;
    QPIC_DYN_SMEM(double2, sm);
    const int nz = g.nz, ny = g.ny;
    double2 * a = sm;
    double2 * w = sm + spb * nz;
    const int row0 = blockIdx.x * spb;
    for (int b = 0; b < spb; b++) {
        const int j = row0 + b;
        if (j >= ny) {
            for (int k = threadIdx.x; k < nz; k += blockDim.x) a[b * nz + k] = make_double2(0.0, 0.0);
            continue;
        }
        const int jp = wrap_inc(j, ny);
        const int jq = (j == ny - 1) ? 0 : j;
        for (int k = threadIdx.x; k < nz; k += blockDim.x) {
            const int t = j * nz + k;
            const int kp = wrap_inc(k, nz);
            const double djz_dxi = (jz[t] - jz_next[t]) * g.inv_dx;
            const double djy_dxi = (jy[t] - jy_next[t]) * g.inv_dx;
            const double jxt = jx[t];
            a[b * nz + k] = make_double2(-D * by[t] - (jx[j * nz + kp] - jxt) * g.inv_dz + djz_dxi,
                                         -D * bz[jq * nz + k] + (jx[jp * nz + k] - jxt) * g.inv_dy - djy_dxi);
        }
    }
    __syncthreads();
    const double2 * z = fft_smem<false, POW2>(a, w, spb, Pz, twz);
    for (int b = 0; b < spb; b++) {
        if (row0 + b >= ny) break;
        for (int k = threadIdx.x; k < nz; k += blockDim.x) spec[(size_t)(row0 + b) * nz + k] = z[b * nz + k];
    }
}

// backward rows of the B solve: by(i) = re/ny/nz, bz(i) = im/ny/nz and the stop test (src/System_3d.cpp:541-552, :569-580)
template <bool POW2>
__global__ void k_fft_rows_b_backward(Ctl * ctl, FftPlan Pz, const double2 * __restrict__ twz, Geom g, const double2 * __restrict__ spec,
                                      double * __restrict__ by, double * __restrict__ bz, double threshold, int iteration, int spb, double inv_scale) {
    QPIC_KERNEL_BEGIN(KID_k_fft_rows_b_backward);
    // blocks of THIS launch may already have raised the flag; only an earlier launch's flag cancels the work
    const int my_tag = ctl->slice * STOP_TAG_STRIDE + iteration + 1;
    const int tag = ctl->stop_tag;
    if ((tag != 0 && tag != my_tag) || ctl->fast_fail) return;
    QPIC_DYN_SMEM(double2, sm);
    const int nz = g.nz, ny = g.ny;
    double2 * a = sm;
    double2 * w = sm + spb * nz;
    const int row0 = blockIdx.x * spb;
    for (int b = 0; b < spb; b++) {
        const bool live = row0 + b < ny;
        for (int k = threadIdx.x; k < nz; k += blockDim.x) a[b * nz + k] = live ? spec[(size_t)(row0 + b) * nz + k] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    const double2 * z = fft_smem<true, POW2>(a, w, spb, Pz, twz);
    bool bad = false;
    for (int b = 0; b < spb; b++) {
        if (row0 + b >= ny) break;
        for (int k = threadIdx.x; k < nz; k += blockDim.x) {
            const double2 v = z[b * nz + k];
            const double vy = scale_result(v.x, ny, nz, inv_scale);
            const double vz = scale_result(v.y, ny, nz, inv_scale);
            by[(size_t)(row0 + b) * nz + k] = vy;
            bz[(size_t)(row0 + b) * nz + k] = vz;
            // !(|v| <= threshold) is true for |v| > threshold and for NaN
            if (!(fabs(vy) <= threshold) || !(fabs(vz) <= threshold)) bad = true;
        }
    }
    if (bad) atomicExch(&ctl->stop_tag, my_tag);
}

// generic complex rows (used by the exposed Helmholtz-pair building block and the tests)
template <bool POW2>
__global__ void k_fft_rows_c2c(FftPlan Pz, const double2 * __restrict__ twz, const double * __restrict__ re, const double * __restrict__ im,
                               double2 * __restrict__ spec, int ny, int nz, int spb) {
    QPIC_KERNEL_BEGIN(KID_k_fft_rows_c2c);
    QPIC_DYN_SMEM(double2, sm);
    double2 * a = sm;
    double2 * w = sm + spb * nz;
    const int row0 = blockIdx.x * spb;
    for (int b = 0; b < spb; b++) {
        const bool live = row0 + b < ny;
        const size_t g = (size_t)(row0 + b) * nz;
        for (int k = threadIdx.x; k < nz; k += blockDim.x) a[b * nz + k] = live ? make_double2(re[g + k], im[g + k]) : make_double2(0.0, 0.0);
    }
    __syncthreads();
    const double2 * z = fft_smem<false, POW2>(a, w, spb, Pz, twz);
    for (int b = 0; b < spb; b++) {
        if (row0 + b >= ny) break;
        for (int k = threadIdx.x; k < nz; k += blockDim.x) spec[(size_t)(row0 + b) * nz + k] = z[b * nz + k];
    }
}

}  // namespace qpic
