// Register-resident transforms for power-of-two lengths (512, 1024, 4096: the BASELINE transverse grids).
//
// Every thread owns 8 points of its sequence, t + (n/8) e for e = 0..7, in registers.  A transform is a chain of radix-8
// Stockham passes (plus one radix-4 / radix-2 pass to finish): butterfly in registers, twiddle, exchange through
// shared memory (two alternating buffers -> one barrier per exchange), next pass.  The ownership pattern is the same
// before the first and after the last pass, so forward transform -> spectral division -> backward transform of a column
// runs without the data ever leaving registers except for the 2 (n = 512) exchanges per transform, and the row kernels
// read / write global memory directly from / to registers, fully coalesced.
//
// Per pass a thread fetches 3 table twiddles (w, w^2, w^4; exact to 1/2 ulp) and forms w^3, w^5, w^6, w^7 with four complex
// products instead of issuing seven dependent table loads.
//
// The spectral division multiplies by a host-built table of 1 / (-D + (2cos(2 pi j/ny)-2)/dy^2 + (2cos(2 pi k/nz)-2)/dz^2)
// (src/System_3d.cpp:699; entry (0,0) = 0, :688-689), stored column-major so that a column reads it contiguously.
#pragma once

#include "qpic_fft.cuh"

namespace qpic {

// Exchange buffers are PADDED: point i of a sequence lives at p2_pad(i) = i + i/8 + i/64.  A radix-8 pass stores the 8 outputs of
// thread t at 8t + e (first pass) / q + 64 p' + 8e (second pass): unpadded, the lanes of a warp sit 128 bytes apart and every
// 16-byte store hits the same banks (32-way / 4-way conflicts; with 4 warps per SM the two exchanges of a 512-point transform spent
// ~0.8 us in the shared-memory pipe).  With the padding both the strided stores and the unit-stride loads touch every bank group
// once per quarter warp.
__host__ __device__ __forceinline__ int p2_pad(int i) { return i + (i >> 3) + (i >> 6); }
__host__ __device__ __forceinline__ int p2_pad_len(int n) { return n + (n >> 3) + (n >> 6) + 1; }

// one transform of the sequence whose exchange buffers are xa / xb (p2_pad_len(n) points each, element i at [p2_pad(i) * stride]:
// stride > 1 interleaves the sequences of a block, see p2_lane_cols); v[e] = point t + (n/8) e
template <bool INV, bool SINGLE = false>
__device__ __forceinline__ void fft_pow2_regs(double2 (&v)[8], double2 * xa, double2 * xb, int t, int log2n, const double2 * __restrict__ tw, int stride = 1) {
    const int le8 = log2n - 3;
    int log2s = 0;
    int rem = log2n;
    int flip = 0;
    while (true) {
        const int lr = rem >= 3 ? 3 : rem;
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
        rem -= lr;
        if (rem == 0) break;  // the finishing pass has p = 0 (no twiddles) and leaves v[e] = point t + (n/8) e
        // a radix-8 pass that is not the last one: twiddle by w_p^k and exchange
        const int p = t >> log2s, q = t & ((1 << log2s) - 1);
        const double2 w1 = twiddle<INV>(tw, p << log2s);
        const double2 w2 = twiddle<INV>(tw, (2 * p) << log2s);
        const double2 w4 = twiddle<INV>(tw, (4 * p) << log2s);
        const double2 w3 = cmul(w1, w2), w5 = cmul(w4, w1), w6 = cmul(w4, w2);
        const double2 w7 = cmul(w4, w3);
        double2 * y = flip ? xb : xa;
        // SINGLE: one buffer instead of two alternating ones (xa == xb, the wide / deep builds of the kernels: half the shared memory, so that more sequences
        // fit a block on the long axes): every thread must have finished reading the previous exchange before it is overwritten
        if (SINGLE && log2s > 0) __syncthreads();
        const int base = q + ((8 * p) << log2s);
        y[p2_pad(base) * stride] = v[0];
        y[p2_pad(base + (1 << log2s)) * stride] = cmul(v[1], w1);
        y[p2_pad(base + (2 << log2s)) * stride] = cmul(v[2], w2);
        y[p2_pad(base + (3 << log2s)) * stride] = cmul(v[3], w3);
        y[p2_pad(base + (4 << log2s)) * stride] = cmul(v[4], w4);
        y[p2_pad(base + (5 << log2s)) * stride] = cmul(v[5], w5);
        y[p2_pad(base + (6 << log2s)) * stride] = cmul(v[6], w6);
        y[p2_pad(base + (7 << log2s)) * stride] = cmul(v[7], w7);
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = y[p2_pad(t + (e << le8)) * stride];
        flip ^= 1;
        log2s += 3;
    }
}

// Where element (global row j, column k) of a spectrum lives.  Single GPU: one buffer spec[j][k] with leading dimension
// cpr = number of columns.  Slab-decomposed runs (qpic_dist.cuh): the columns are dealt out in blocks of cpr to the ranks;
// base[q] is rank q's column buffer [ny][cpr] (a peer pointer over NVLink), so a forward row kernel's store IS the transpose.
struct SpecMap {
    double2 * base[8];
    int cpr;
    int world;
};

// (the peer table is only ever indexed with compile-time constants: a dynamically indexed kernel parameter would be copied
// to local memory by every thread)
__device__ __forceinline__ double2 * spec_base(const SpecMap & m, int q) {
    double2 * b = m.base[0];
#pragma unroll
    for (int i = 1; i < 8; i++) {
        if (q == i) b = m.base[i];
    }
    return b;
}

// DIST is a template parameter of the row kernels so that the single-GPU instantiation keeps plain [j][k] indexing
template <bool DIST>
__device__ __forceinline__ double2 * spec_at(const SpecMap & m, int j, int k) {
    if (!DIST) return m.base[0] + (size_t)j * m.cpr + k;
    const int q = k / m.cpr;
    return spec_base(m, q) + (size_t)j * m.cpr + (k - q * m.cpr);
}

// thread layout shared by all kernels below: blockDim.x = spb * n/8; sequence b = threadIdx.x / (n/8), lane t
// Every kernel below exists in three builds (last template argument MODE):
//   0 narrow  <= 512 threads per block (no register cap: the 512-point axes of config #2 run 128-thread blocks), two alternating exchange buffers;
//   1 wide    <= 896 threads = 7 sequences of a 1024-point axis in one block (1024 sequences on 148 SMs = one wave of 147 blocks) at 72 registers
//             (a few spilled words) with ONE exchange buffer per sequence (one more barrier per exchange);
//   2 deep    <= 512 threads at 64 registers with one exchange buffer: TWO blocks per SM where a sequence fills a block (4096-point axes: 512
//             threads and 66 KB per sequence), so that one block's loads overlap the other's passes.
// size_fft_launch picks the build that needs the fewest waves (ties: narrow, wide, deep).
constexpr int P2_WIDE_THREADS = 896, P2_NARROW_THREADS = 512;
#define P2_BOUNDS __launch_bounds__(MODE == 1 ? P2_WIDE_THREADS : P2_NARROW_THREADS, MODE == 2 ? 2 : 0)
struct Pow2Lane {
    int b, t, e8;
    double2 *xa, *xb;
};

// SINGLE: one exchange buffer per sequence (see fft_pow2_regs)
template <bool SINGLE = false>
__device__ __forceinline__ Pow2Lane pow2_lane(double2 * sm, int log2n) {
    Pow2Lane L;
    const int le8 = log2n - 3;
    const int plen = p2_pad_len(1 << log2n);
    L.e8 = 1 << le8;
    L.b = threadIdx.x >> le8;
    L.t = threadIdx.x & (L.e8 - 1);
    if (SINGLE) {
        L.xa = sm + (size_t)L.b * plen;
        L.xb = L.xa;
    } else {
        L.xa = sm + (size_t)L.b * 2 * plen;
        L.xb = L.xa + plen;
    }
    return L;
}

// ------------------------------------------------------------------------------------------------ real field (psi)

// forward rows, two real rows per sequence -> half spectra spec[ny][nzh]
// rows [row0, row0 + ny) of src (a full-size array)
template <bool DIST, int MODE = 0>
__global__ void P2_BOUNDS k_p2_rows_r2c(const Ctl * ctl, int log2n, const double2 * __restrict__ twz, const double * __restrict__ src, SpecMap spec, int ny,
                              int row0) {
    QPIC_KERNEL_BEGIN(KID_k_p2_rows_r2c);
    QPIC_DYN_SMEM(double2, sm);
    const Pow2Lane L = pow2_lane<MODE != 0>(sm, log2n);
    const int nz = 1 << log2n, nzh = nz / 2 + 1;
    const int j0 = 2 * (blockIdx.x * (blockDim.x >> (log2n - 3)) + L.b), j1 = j0 + 1;
    double2 v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int k = L.t + e * L.e8;
        v[e] = make_double2(j0 < ny ? src[(size_t)(row0 + j0) * nz + k] : 0.0, j1 < ny ? src[(size_t)(row0 + j1) * nz + k] : 0.0);
    }
    if (cancelled(ctl)) return;  // tested after the loads are in flight (uniform over the block)
    fft_pow2_regs<false, MODE != 0>(v, L.xa, L.xb, L.t, log2n, twz);
    QPIC_KERNEL_TAIL();
    // the split needs Z_k and Z_{n-k}: one more exchange.  The last exchange of the transform used the other buffer
    // (or none at all), so this buffer is free once every thread has left the transform's final read.
    __syncthreads();
    double2 * z = L.xa;
#pragma unroll
    for (int e = 0; e < 8; e++) z[p2_pad(L.t + e * L.e8)] = v[e];
    __syncthreads();
    if (j0 >= ny) return;
    for (int k = L.t; k < nzh; k += L.e8) {
        const double2 zk = z[p2_pad(k)];
        const double2 zm = z[p2_pad(k == 0 ? 0 : nz - k)];
        // U0 = (Z_k + conj(Z_{n-k})) / 2,  U1 = (Z_k - conj(Z_{n-k})) / (2i)
        *spec_at<DIST>(spec, row0 + j0, k) = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
        if (j1 < ny) *spec_at<DIST>(spec, row0 + j1, k) = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));
    }
}

// backward rows: two half spectra -> one complex sequence -> two real rows, scaled by 1/ny/nz (src/System_3d.cpp:361).
// With fix != 0 the potential fix-up is fused in: psi = max(u - mean(first column U first row of u), floor)
// (normalize_potential + increase_minimum, src/System_3d.cpp:30-47, :830-841).  The border mean is obtained from the
// spectrum U = ny nz FFT2(u) that the column kernel leaves behind, so no second pass over the field is needed:
//     sum_j u(j,0)  = (1/nz) sum_kz U(0,kz)            (row0[], Hermitian in kz)
//     sum_k u(0,k)  = (1/ny) sum_ky U(ky,0)            (colsum[0])
//     u(0,0)        = (1/(ny nz)) sum_ky sum_kz U      (colsum[], Hermitian in kz)
// The column kernel's last block reduces these sums with a fixed-order tree and leaves the mean in *border_mean.
// `ny` rows starting at global row `row0` of a field with ny_total rows.  The border mean is assembled from the per-rank
// partial sums the column kernels left in partials[rank][0..2] (one rank on a single GPU), in rank order.
template <bool DIST, int MODE = 0>
__global__ void P2_BOUNDS k_p2_rows_c2r(const Ctl * ctl, int log2n, const double2 * __restrict__ twz, SpecMap spec, double * __restrict__ dst, int ny, int row0,
                              int ny_total, double inv_scale, int fix, const double * __restrict__ partials, double floor_value,
                              const double2 * __restrict__ red_row0, const double2 * __restrict__ red_colsum) {
    QPIC_KERNEL_BEGIN(KID_k_p2_rows_c2r);
    QPIC_DYN_SMEM(double2, sm);
    const Pow2Lane L = pow2_lane<MODE != 0>(sm, log2n);
    const int nz = 1 << log2n;
    const int j0 = 2 * (blockIdx.x * (blockDim.x >> (log2n - 3)) + L.b), j1 = j0 + 1;
    // loads now, arithmetic after the transform (the divisions then hide behind the spectrum loads)
    double pa = 0.0, pb = 0.0, pc = 0.0;
    if (fix && !red_row0) {
        pc = partials[2];  // column 0 belongs to rank 0
#pragma unroll
        for (int r = 0; r < 8; r++) {
            if (r < spec.world) {
                pa += partials[4 * r + 0];
                pb += partials[4 * r + 1];
            }
        }
    }
    double2 v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int i = L.t + e * L.e8;
        const bool mirror = i > nz / 2;
        const int k = mirror ? nz - i : i;
        double2 v0 = (j0 < ny) ? *spec_at<DIST>(spec, row0 + j0, k) : make_double2(0.0, 0.0);
        double2 v1 = (j1 < ny) ? *spec_at<DIST>(spec, row0 + j1, k) : make_double2(0.0, 0.0);
        if (k == 0 || 2 * k == nz) {
            // c2r ignores the imaginary part of the DC / Nyquist bins (they are rounding noise for a real field)
            v0.y = 0.0;
            v1.y = 0.0;
        }
        // Z_k = V0 + i V1;  Z_{n-k} = conj(V0) + i conj(V1)
        v[e] = mirror ? make_double2(v0.x + v1.y, v1.x - v0.y) : make_double2(v0.x - v1.y, v0.y + v1.x);
    }
    if (fix && red_row0) {
        // single GPU: the Hermitian-weighted sums of Re U(0,k) and Re sum_j U(j,k) over the half spectrum, straight from what the
        // column kernel left per column.  Every warp does the whole (tiny) reduction in the same fixed order -- lane l takes columns
        // l, l + 32, ... ascending, then an xor-shuffle tree -- so all warps of all blocks hold identical sums and no barrier is
        // needed; the loads sit behind the spectrum loads above and overlap them.
        // (The last warp of a block may be partial -- blockDim is a multiple of 8 only -- so the tree runs over the first W lanes,
        // W = the largest power of two that every warp of the block has, and lane 0 broadcasts.)
        const int lane = threadIdx.x & 31;
        const int rem = blockDim.x & 31;
        const int W = rem == 0 ? 32 : (rem >= 16 ? 16 : 8);
        if (W == 32) {
            for (int kk = lane; kk <= nz / 2; kk += 32) {
                const double wgt = (kk == 0 || 2 * kk == nz) ? 1.0 : 2.0;
                pa += wgt * red_row0[kk].x;
                pb += wgt * red_colsum[kk].x;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                pa += __shfl_xor_sync(0xffffffffu, pa, o);
                pb += __shfl_xor_sync(0xffffffffu, pb, o);
            }
        } else {
            const unsigned mask = __activemask();
            if (lane < W) {
                for (int kk = lane; kk <= nz / 2; kk += W) {
                    const double wgt = (kk == 0 || 2 * kk == nz) ? 1.0 : 2.0;
                    pa += wgt * red_row0[kk].x;
                    pb += wgt * red_colsum[kk].x;
                }
            }
            for (int o = W >> 1; o > 0; o >>= 1) {
                pa += __shfl_xor_sync(mask, pa, o);
                pb += __shfl_xor_sync(mask, pb, o);
            }
            pa = __shfl_sync(mask, pa, 0);
            pb = __shfl_sync(mask, pb, 0);
        }
        pc = red_colsum[0].x;
    }
    if (cancelled(ctl)) return;
    fft_pow2_regs<true, MODE != 0>(v, L.xa, L.xb, L.t, log2n, twz);
    QPIC_KERNEL_TAIL();
    double mean = 0.0;
    if (fix) {
        const double s_col0 = pa / nz;              // sum_j u(j,0)
        const double s_row0 = pc / ny_total;        // sum_k u(0,k)
        const double u00 = pb / ny_total / nz;      // u(0,0), counted in both
        mean = (s_col0 + s_row0 - u00) / (ny_total + nz - 1);
    }
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int k = L.t + e * L.e8;
        double a = scale_result(v[e].x, ny_total, nz, inv_scale), b = scale_result(v[e].y, ny_total, nz, inv_scale);
        if (fix) {
            a -= mean;
            b -= mean;
            if (a < floor_value) a = floor_value;
            if (b < floor_value) b = floor_value;
        }
        if (j0 < ny) dst[(size_t)(row0 + j0) * nz + k] = a;
        if (j1 < ny) dst[(size_t)(row0 + j1) * nz + k] = b;
    }
}

// ------------------------------------------------------------------------------------------------ columns + symbol
// one column per sequence; inv_symbol is [ncols][ny] (column-major for this kernel), entry (k=0, j=0) = 0
template <int MODE = 0>
__global__ void P2_BOUNDS k_p2_cols_solve(const Ctl * ctl, int log2n, const double2 * __restrict__ twy, double2 * __restrict__ spec, int ld, int ncols,
                                const double * __restrict__ inv_symbol, double2 * __restrict__ row0, double2 * __restrict__ colsum,
                                int * __restrict__ ticket, SpecMap partials, int my_rank, int col0, int nz) {
    QPIC_KERNEL_BEGIN(colsum ? KID_k_p2_cols_solve : KID_COUNT + 16);
    QPIC_DYN_SMEM(double2, sm);
    // The spb columns of a block are ADJACENT and interleaved over the threads (column index fastest) and in shared memory, so the lanes
    // of a warp touch spb * 16 contiguous bytes of every spectrum row: full 32-byte sectors from two columns per block on (one
    // column per 64 consecutive threads used half of every sector and 32 lines per warp-wide load).
    Pow2Lane L;
    const int spb = blockDim.x >> (log2n - 3);
    L.e8 = 1 << (log2n - 3);
    L.b = threadIdx.x % spb;
    L.t = threadIdx.x / spb;
    L.xa = sm + L.b;
    L.xb = MODE != 0 ? L.xa : L.xa + (size_t)p2_pad_len(1 << log2n) * spb;
    const int k = blockIdx.x * spb + L.b;
    const bool live = k < ncols;
    const int ny = 1 << log2n;
    double2 v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = live ? spec[(size_t)(L.t + e * L.e8) * ld + k] : make_double2(0.0, 0.0);
    // 1 / symbol: fetched together with the column so that it is there when the forward transform ends
    double rinv[8];
#pragma unroll
    for (int e = 0; e < 8; e++) rinv[e] = live ? inv_symbol[(size_t)k * ny + L.t + e * L.e8] : 0.0;
    if (cancelled(ctl)) return;
    fft_pow2_regs<false, MODE != 0>(v, L.xa, L.xb, L.t, log2n, twy, spb);
    if (live) {
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int j = L.t + e * L.e8;
            v[e] = (j == 0 && k + col0 == 0) ? make_double2(0.0, 0.0) : make_double2(v[e].x * rinv[e], v[e].y * rinv[e]);
        }
    }
    // what the fused potential fix-up of k_p2_rows_c2r needs per column: U(0,k) -- element 0 before the backward transform -- and
    // sum_j U(j,k), which IS element 0 of the unnormalised backward transform, so no reduction is needed: thread t = 0 of the
    // sequence owns index 0 in v[0] both before and after the transform
    if (colsum && live && L.t == 0) row0[k] = v[0];
    // the forward transform's last exchange buffer may still be read by slower threads of this sequence
    __syncthreads();
    fft_pow2_regs<true, MODE != 0>(v, L.xa, L.xb, L.t, log2n, twy, spb);
    QPIC_KERNEL_TAIL();
    if (live) {
#pragma unroll
        for (int e = 0; e < 8; e++) spec[(size_t)(L.t + e * L.e8) * ld + k] = v[e];
        if (colsum && L.t == 0) colsum[k] = v[0];
    }
    if (colsum && ticket) {
        // (Slab-decomposed sweeps only; on a single GPU the backward row kernel reduces row0[] / colsum[] itself, which removes
        // this serial tail -- fence, ticket, one block's reduction -- from every real solve: 0.763 -> 0.751 ms/slice.)
        // The block that takes the last ticket sees every block's row0[] / colsum[] (release: fence before the ticket;
        // acquire: fence after it) and reduces them to the border mean -- see k_p2_rows_c2r for the formulas.
        __shared__ int is_last;
        __shared__ double red_a[32], red_b[32];  // one partial per warp
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
        __syncthreads();
        if (is_last) {
            __threadfence();
            double pa = 0.0, pb = 0.0;  // Hermitian-weighted sums of Re U(0,k) and Re sum_j U(j,k) over this rank's columns
            for (int kk = threadIdx.x; kk < ncols; kk += blockDim.x) {
                const int kg = kk + col0;
                const double wgt = (kg == 0 || 2 * kg == nz) ? 1.0 : 2.0;
                pa += wgt * row0[kk].x;
                pb += wgt * colsum[kk].x;
            }
            // fixed-order reduction: shuffle tree inside each warp, then the warps' partials in order
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                pa += __shfl_down_sync(0xffffffffu, pa, o);
                pb += __shfl_down_sync(0xffffffffu, pb, o);
            }
            if ((threadIdx.x & 31) == 0) {
                red_a[threadIdx.x >> 5] = pa;
                red_b[threadIdx.x >> 5] = pb;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                const int nwarps = (blockDim.x + 31) >> 5;
                for (int w = 1; w < nwarps; w++) {
                    red_a[0] += red_a[w];
                    red_b[0] += red_b[w];
                }
            }
            if (threadIdx.x == 0) {
                // every rank gets this rank's partial sums (slot my_rank of its table); k_p2_rows_c2r adds them up in rank order
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    if (q < partials.world) {
                        double * dst = reinterpret_cast<double *>(partials.base[q]) + 4 * my_rank;
                        dst[0] = red_a[0];
                        dst[1] = red_b[0];
                        dst[2] = colsum[0].x;  // meaningful on the rank that owns column 0
                    }
                }
                *ticket = 0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ complex field (By + i Bz)

// see k_fft_rows_b_forward (qpic_fft.cuh) for the source terms (src/System_3d.cpp:489-495, :530-536, :555-564 incl. :563)
// PRE: the sources were already formed by the jy/jz deposit (TileJob::b_src) and are simply loaded
template <bool DIST, bool PRE, int MODE = 0>
__global__ void P2_BOUNDS k_p2_rows_b_forward(const Ctl * ctl, int log2n, const double2 * __restrict__ twz, Geom g, double D, const double * __restrict__ by,
                                    const double * __restrict__ bz, const double * __restrict__ jx, const double * __restrict__ jy,
                                    const double * __restrict__ jz, const double * __restrict__ jy_next, const double * __restrict__ jz_next,
                                    const double2 * __restrict__ b_src, SpecMap spec, int row0, int nrows) {
    QPIC_KERNEL_BEGIN(KID_k_p2_rows_b_forward);
    QPIC_DYN_SMEM(double2, sm);
    const Pow2Lane L = pow2_lane<MODE != 0>(sm, log2n);
    const int nz = g.nz, ny = g.ny;
    const int jl = blockIdx.x * (blockDim.x >> (log2n - 3)) + L.b;
    const int j = row0 + jl;
    const bool live = jl < nrows;
    const int jp = live ? wrap_inc(j, ny) : 0;
    const int jq = (j == ny - 1) ? 0 : j;
    double2 v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int k = L.t + e * L.e8;
        if (live && PRE) {
            v[e] = b_src[j * nz + k];
        } else if (live) {
            const int t = j * nz + k;
            const int kp = wrap_inc(k, nz);
            const double djz_dxi = (jz[t] - jz_next[t]) * g.inv_dx;
            const double djy_dxi = (jy[t] - jy_next[t]) * g.inv_dx;
            const double jxt = jx[t];
            v[e] = make_double2(-D * by[t] - (jx[j * nz + kp] - jxt) * g.inv_dz + djz_dxi, -D * bz[jq * nz + k] + (jx[jp * nz + k] - jxt) * g.inv_dy - djy_dxi);
        } else {
            v[e] = make_double2(0.0, 0.0);
        }
    }
    if (cancelled(ctl)) return;
    fft_pow2_regs<false, MODE != 0>(v, L.xa, L.xb, L.t, log2n, twz);
    QPIC_KERNEL_TAIL();
    if (live) {
#pragma unroll
        for (int e = 0; e < 8; e++) *spec_at<DIST>(spec, j, L.t + e * L.e8) = v[e];
    }
}

// by(i) = re/ny/nz, bz(i) = im/ny/nz and the stop test (src/System_3d.cpp:541-552, :569-580)
template <bool DIST, int MODE = 0>
__global__ void P2_BOUNDS k_p2_rows_b_backward(Ctl * ctl, int log2n, const double2 * __restrict__ twz, Geom g, SpecMap spec, double * __restrict__ by,
                                     double * __restrict__ bz, double threshold, int iteration, double inv_scale, int row0, int nrows) {
    QPIC_KERNEL_BEGIN(KID_k_p2_rows_b_backward);
    // blocks of THIS launch may already have raised the flag; only an earlier launch's flag cancels the work
    QPIC_DYN_SMEM(double2, sm);
    const Pow2Lane L = pow2_lane<MODE != 0>(sm, log2n);
    const int nz = g.nz, ny = g.ny;
    const int jl = blockIdx.x * (blockDim.x >> (log2n - 3)) + L.b;
    const int j = row0 + jl;
    const bool live = jl < nrows;
    double2 v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = live ? *spec_at<DIST>(spec, j, L.t + e * L.e8) : make_double2(0.0, 0.0);
    const int my_tag = ctl->slice * STOP_TAG_STRIDE + iteration + 1;
    const int tag = ctl->stop_tag;
    if ((tag != 0 && tag != my_tag) || ctl->fast_fail) return;
    fft_pow2_regs<true, MODE != 0>(v, L.xa, L.xb, L.t, log2n, twz);
    QPIC_KERNEL_TAIL();
    if (!live) return;
    bool bad = false;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const int k = L.t + e * L.e8;
        const double vy = scale_result(v[e].x, ny, nz, inv_scale);
        const double vz = scale_result(v[e].y, ny, nz, inv_scale);
        by[(size_t)j * nz + k] = vy;
        bz[(size_t)j * nz + k] = vz;
        // !(|v| <= threshold) is true for |v| > threshold and for NaN
        if (!(fabs(vy) <= threshold) || !(fabs(vz) <= threshold)) bad = true;
    }
    if (bad) atomicExch(&ctl->stop_tag, my_tag);
}

// plain complex rows (exposed Helmholtz-pair building block)
template <int MODE = 0>
__global__ void P2_BOUNDS k_p2_rows_c2c(int log2n, const double2 * __restrict__ twz, const double * __restrict__ re, const double * __restrict__ im,
                              double2 * __restrict__ spec, int ny) {
    QPIC_KERNEL_BEGIN(KID_k_p2_rows_c2c);
    QPIC_DYN_SMEM(double2, sm);
    const Pow2Lane L = pow2_lane<MODE != 0>(sm, log2n);
    const int nz = 1 << log2n;
    const int j = blockIdx.x * (blockDim.x >> (log2n - 3)) + L.b;
    const bool live = j < ny;
    double2 v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const size_t gidx = (size_t)j * nz + L.t + e * L.e8;
        v[e] = live ? make_double2(re[gidx], im[gidx]) : make_double2(0.0, 0.0);
    }
    fft_pow2_regs<false, MODE != 0>(v, L.xa, L.xb, L.t, log2n, twz);
    if (live) {
#pragma unroll
        for (int e = 0; e < 8; e++) spec[(size_t)j * nz + L.t + e * L.e8] = v[e];
    }
}

}  // namespace qpic
