// Beam ("bunch") macroparticles on the device: the part of System_3d::run around solve_wakefield
// (reference src/System_3d.cpp:199-236) -- deposit_bunch_densities (:926-946), deposit_wakefield_to_bunch (:843-860) and
// advance_bunch_particles with qed == false (:862-886), with the 3-D deposit / gather primitives of src/array_utils.cpp:290-375.
//
// Data layout: bunch_particle_3d (src/containers.h:45-61) as SoA of doubles + an int species id; rho_bunch, jx_bunch and the
// five field arrays are the resident (nx, ny, nz) arrays of the sweep, so nothing crosses PCIe between time iterations.
//
// Deterministic deposit, no floating-point atomics: the particles are counting-sorted by the cell (i1, j1, k1) of their lower
// corner (integer atomics; every bucket is put into ascending particle order), then ONE THREAD PER NODE adds the
// contributions of the particles in the 8 cells that touch it, cells and particles in a fixed order -- and writes the node
// once, which also replaces the reference's zeroing pass (:927-935).  All HBM-bound FP64 / integer work.
#pragma once

#include "qpic_kernels.cuh"

namespace qpic {

constexpr int BUNCH_MAX_SPECIES = 16;

struct BunchSpecies {  // Species (src/Species_register.h:6-15) by species id
    double charge[BUNCH_MAX_SPECIES];
    double charge_to_mass[BUNCH_MAX_SPECIES];
    int is_photon[BUNCH_MAX_SPECIES];
};

struct Bunch {
    int np;
    double *x, *y, *z, *px, *py, *pz, *gamma, *n, *ex, *ey, *ez, *by, *bz, *chi;
    int * species;
};

// floor(t) as an index on a periodic axis of n points, and the fraction t - floor(t).  The reference wraps only
// -1 (j1 += n) and n - 1 (j2 = (j1 + 1) % n), i.e. it is defined for -d <= y < l (src/array_utils.cpp:304-316); this is the
// periodic continuation of that.
__device__ __forceinline__ void locate_periodic(double t, int n, int & j1, double & f) {
    const double fl = floor(t);
    f = t - fl;
    long long j = (long long)fl % n;
    if (j < 0) j += n;
    j1 = (int)j;
}

// cell of the particle's lower corner on the node-centred grid of rho_bunch / jx_bunch; -1 = outside [0, nx - 1) in x,
// where the reference's deposit returns without touching the array (src/array_utils.cpp:299-301)
__global__ void k_bunch_locate_count(Bunch B, Geom g, int * __restrict__ cell, int * __restrict__ count) {
    QPIC_KERNEL_BEGIN(KID_k_bunch_locate_count);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.np) return;
    const double fl = floor(B.x[p] / g.dx);
    if (!(fl >= 0.0 && fl + 1.0 < (double)g.nx)) {
        cell[p] = -1;
        return;
    }
    int j1, k1;
    double f;
    locate_periodic(B.y[p] / g.dy, g.ny, j1, f);
    locate_periodic(B.z[p] / g.dz, g.nz, k1, f);
    const int c = ((int)fl * g.ny + j1) * g.nz + k1;
    cell[p] = c;
    atomicAdd(&count[c], 1);
}

// count[] -> start[] (exclusive prefix sum over all nx*ny*nz cells), count[] <- start[] (it becomes the fill cursor)
__global__ void __launch_bounds__(SCAN_THREADS)
    k_bunch_scan_apply(int * __restrict__ count, int n, const int * __restrict__ sums, int * __restrict__ start) {
    QPIC_KERNEL_BEGIN(KID_k_bunch_scan_apply);
    const int chunk_offset = sums[blockIdx.x];
    const long long base = (long long)blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    int c[SCAN_ITEMS];
    int v = 0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; t++) {
        c[t] = (base + t < n) ? count[base + t] : 0;
        v += c[t];
    }
    int total;
    int run = chunk_offset + block_exclusive_scan(v, total);
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; t++) {
        if (base + t < n) {
            start[base + t] = run;
            count[base + t] = run;
            run += c[t];
        }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) start[n] = run;
}

// bucket fill in whatever order the integer atomics resolve ...
__global__ void k_bunch_fill(int np, const int * __restrict__ cell, int * __restrict__ cursor, int * __restrict__ tmp) {
    QPIC_KERNEL_BEGIN(KID_k_bunch_fill);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const int c = cell[p];
    if (c < 0) return;
    tmp[atomicAdd(&cursor[c], 1)] = p;
}

// ... then every particle takes the place of its rank inside the bucket: ascending particle index, the order in which a
// single-threaded reference adds them
__global__ void k_bunch_rank(int np, const int * __restrict__ cell, const int * __restrict__ start, const int * __restrict__ tmp,
                             int * __restrict__ perm) {
    QPIC_KERNEL_BEGIN(KID_k_bunch_rank);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    const int c = cell[p];
    if (c < 0) return;
    const int b = start[c], e = start[c + 1];
    int rank = 0;
    for (int t = b; t < e; t++) rank += (tmp[t] < p) ? 1 : 0;
    perm[b + rank] = p;
}

// rho_bunch(i,j,k), jx_bunch(i,j,k) = sum over the particles of cells (i - di, j - dj, k - dk), di, dj, dk in {0, 1}, of
// value * wx * wy * wz with w = 1 - frac (d = 0) or frac (d = 1): the transpose of the 8-point scatter of
// src/array_utils.cpp:323-338, every product evaluated left to right as there; values as in src/System_3d.cpp:942-944
__global__ void k_bunch_deposit(Geom g, Bunch B, BunchSpecies S, const int * __restrict__ start, const int * __restrict__ perm,
                                double * __restrict__ rho_bunch, double * __restrict__ jx_bunch) {
    QPIC_KERNEL_BEGIN(KID_k_bunch_deposit);
    const size_t total = (size_t)g.nx * g.n2;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(t % g.nz);
        const int j = (int)((t / g.nz) % g.ny);
        const int i = (int)(t / g.n2);
        double rho = 0.0, jx = 0.0;
#pragma unroll
        for (int di = 0; di < 2; di++) {
            const int ic = i - di;
            if (ic < 0 || ic > g.nx - 2) continue;  // cells exist for i1 in [0, nx - 2]
#pragma unroll
            for (int dj = 0; dj < 2; dj++) {
                const int jc = dj ? wrap_dec(j, g.ny) : j;
#pragma unroll
                for (int dk = 0; dk < 2; dk++) {
                    const int kc = dk ? wrap_dec(k, g.nz) : k;
                    const int c = (ic * g.ny + jc) * g.nz + kc;
                    const int b = start[c], e = start[c + 1];
                    for (int s = b; s < e; s++) {
                        const int p = perm[s];
                        const double tx = B.x[p] / g.dx, ty = B.y[p] / g.dy, tz = B.z[p] / g.dz;
                        const double fx = tx - floor(tx), fy = ty - floor(ty), fz = tz - floor(tz);
                        const double wx = di ? fx : 1 - fx, wy = dj ? fy : 1 - fy, wz = dk ? fz : 1 - fz;
                        const double charge_density = S.charge[B.species[p]] * B.n[p];
                        rho += charge_density * wx * wy * wz;
                        jx += charge_density * B.px[p] / B.gamma[p] * wx * wy * wz;
                    }
                }
            }
        }
        rho_bunch[t] = rho;
        jx_bunch[t] = jx;
    }
}

// array_to_particle(x, y, z, array3d) for an array with origin (ox, oy, oz): src/array_utils.cpp:340-375
__device__ __forceinline__ double gather3d(const double * __restrict__ a, const Geom & g, double x, double y, double z, double ox, double oy,
                                           double oz) {
    const double tx = (x - ox) / g.dx;
    const double flx = floor(tx);
    if (!(flx >= 0.0 && flx + 1.0 < (double)g.nx)) return 0.0;
    const int i1 = (int)flx;
    const double xf = tx - flx;
    int j1, k1;
    double yf, zf;
    locate_periodic((y - oy) / g.dy, g.ny, j1, yf);
    locate_periodic((z - oz) / g.dz, g.nz, k1, zf);
    const int j2 = wrap_inc(j1, g.ny), k2 = wrap_inc(k1, g.nz);
    const double * lo = a + (size_t)i1 * g.n2;
    const double * hi = lo + g.n2;
    return lo[g.nz * j1 + k1] * (1 - xf) * (1 - yf) * (1 - zf) + lo[g.nz * j2 + k1] * (1 - xf) * yf * (1 - zf) +
           lo[g.nz * j1 + k2] * (1 - xf) * (1 - yf) * zf + lo[g.nz * j2 + k2] * (1 - xf) * yf * zf +
           hi[g.nz * j1 + k1] * xf * (1 - yf) * (1 - zf) + hi[g.nz * j2 + k1] * xf * yf * (1 - zf) + hi[g.nz * j1 + k2] * xf * (1 - yf) * zf +
           hi[g.nz * j2 + k2] * xf * yf * zf;
}

// deposit_wakefield_to_bunch (src/System_3d.cpp:843-860); staggering of the five arrays: src/System_3d.cpp:178-182
__global__ void k_bunch_gather(Bunch B, Geom g, const double * __restrict__ ex, const double * __restrict__ ey, const double * __restrict__ ez,
                               const double * __restrict__ by, const double * __restrict__ bz, double field_schwinger) {
    QPIC_KERNEL_BEGIN(KID_k_bunch_gather);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.np) return;
    const double x = B.x[p], y = B.y[p], z = B.z[p];
    const double pex = gather3d(ex, g, x, y, z, -0.5 * g.dx, 0.0, 0.0);
    const double pey = gather3d(ey, g, x, y, z, 0.0, 0.5 * g.dy, 0.0);
    const double pez = gather3d(ez, g, x, y, z, 0.0, 0.0, 0.5 * g.dz);
    const double pby = gather3d(by, g, x, y, z, 0.0, 0.0, 0.5 * g.dz);
    const double pbz = gather3d(bz, g, x, y, z, 0.0, 0.5 * g.dy, 0.0);
    B.ex[p] = pex;
    B.ey[p] = pey;
    B.ez[p] = pez;
    B.by[p] = pby;
    B.bz[p] = pbz;
    const double gamma = B.gamma[p], px = B.px[p], py = B.py[p], pz = B.pz[p];
    const double fx = gamma * pex + py * pbz - pz * pby;
    const double fy = gamma * pey - px * pbz;
    const double fz = gamma * pez + px * pby;
    const double f_long = px * pex + py * pey + pz * pez;
    B.chi[p] = sqrt(fx * fx + fy * fy + fz * fz - f_long * f_long) / field_schwinger;
}

// advance_bunch_particles with qed == false (src/System_3d.cpp:862-886)
__global__ void k_bunch_advance(Bunch B, BunchSpecies S, double dt) {
    QPIC_KERNEL_BEGIN(KID_k_bunch_advance);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.np) return;
    const int s = B.species[p];
    double px = B.px[p], py = B.py[p], pz = B.pz[p];
    const double gamma = B.gamma[p];
    B.x[p] += dt * (px / gamma - 1);
    B.y[p] += dt * py / gamma;
    B.z[p] += dt * pz / gamma;
    if (S.is_photon[s]) {
        B.gamma[p] = sqrt(px * px + py * py + pz * pz);
    } else {
        const double cmr = S.charge_to_mass[s];
        const double pex = B.ex[p], pey = B.ey[p], pez = B.ez[p], pby = B.by[p], pbz = B.bz[p];
        const double fx = cmr * (pex + (py * pbz - pz * pby) / gamma);
        const double fy = cmr * (pey - px * pbz / gamma);
        const double fz = cmr * (pez + px * pby / gamma);
        px += dt * fx;
        py += dt * fy;
        pz += dt * fz;
        B.px[p] = px;
        B.py[p] = py;
        B.pz[p] = pz;
        B.gamma[p] = sqrt(1 + px * px + py * py + pz * pz);
    }
}

// ------------------------------------------------------------------------------------------------ photon emission (qed = true)

// constants of W_new, formed on the host with the reference's expressions (src/monte_carlo.cpp:17, src/constants.h:3-12)
struct QedConstants {
    double coef;            // m c^2 / (hbar omega)
    double fine_structure;  // e^2 / (hbar c)
};

// powf as the reference's libm evaluates it (single precision, correctly rounded in practice): the double-precision pow of the
// float arguments, rounded to float.  CUDA's own powf differs from glibc's in the last bits, which would move the emission
// probabilities by ~1e-7 relative.
__device__ __forceinline__ double powf_like_libm(double a, double b) { return (double)(float)pow((double)(float)a, (double)(float)b); }

// W_new (src/monte_carlo.cpp:11-41)
__device__ __forceinline__ double w_new(double g, double r, double chi, const QedConstants & Q) {
    const double PI = 3.141592653589793238;
    if ((1.0 - r) * g < 1.0 || chi == 0.0 || r == 0.0) return 0.0;
    double rm = 1.0;
    const bool strong = chi > 0.13333;
    if (!strong) {
        rm = chi / 0.13333;
        r = r * rm;
    }
    const double y = r / ((1.0 - r) * chi);
    const double x = powf_like_libm(y, 2.0 / 3.0);
    const double exp_v = exp(-2.0 * y / 3.0);
    const double iAiryapp = exp_v / (2.0 * sqrt(PI)) * powf_like_libm(x + 0.80049, -0.75);
    const double dAiryapp = -exp_v / (2.0 * sqrt(PI)) * powf_like_libm((x + 0.70861 * (1.0 - 0.65 * x / (1.0 + x * x))), 0.25);
    const double bracket = iAiryapp + (2.0 / x + chi * r * sqrt(x)) * dAiryapp;
    if (strong) return -Q.fine_structure * Q.coef / (g * g) * bracket;
    return -rm * Q.fine_structure * Q.coef / (g * g) * bracket;
}

// Philox4x32-10 (Salmon et al., SC'11): counter = (particle, step, stream, 0), key = seed.  Used when the caller does not supply
// the uniform numbers; every (seed, step, particle) gives the same two numbers whatever the launch shape.
__device__ __forceinline__ void philox4x32_10(unsigned int c[4], unsigned int k0, unsigned int k1) {
#pragma unroll
    for (int round = 0; round < 10; round++) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const unsigned int n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0;
        c[1] = n1;
        c[2] = n2;
        c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ void philox_uniform_pair(unsigned long long seed, unsigned int step, unsigned int particle, double & u0, double & u1) {
    unsigned int c[4] = {particle, step, 0x51454421u, 0u};
    philox4x32_10(c, (unsigned int)seed, (unsigned int)(seed >> 32));
    const unsigned long long a = ((unsigned long long)c[0] << 32) | c[1], b = ((unsigned long long)c[2] << 32) | c[3];
    u0 = (double)(a >> 11) * (1.0 / 9007199254740992.0);  // 53 bits, [0, 1)
    u1 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
}

struct QedScratch {
    int * flag;                      // 1 where the particle emitted this step (then overwritten by its exclusive prefix sum)
    double *qx, *qy, *qz;            // momentum of the emitted photon, by particle
    unsigned long long * max_bits;   // bit pattern of the largest non-negative emission probability (atomicMax: order independent)
};

// advance_bunch_particles with qed == true (src/System_3d.cpp:862-921).  u_emit / u_accept: the two rand_double() values of the
// particle (:890, :897), or NULL = Philox.  Keeps `p.py -= photon.px` (:917).
__global__ void k_bunch_advance_qed(Bunch B, BunchSpecies S, double dt, QedConstants Q, const double * __restrict__ u_emit,
                                    const double * __restrict__ u_accept, unsigned long long seed, unsigned int step, QedScratch X) {
    QPIC_KERNEL_BEGIN(KID_k_bunch_advance_qed);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.np) return;
    const int s = B.species[p];
    double px = B.px[p], py = B.py[p], pz = B.pz[p];
    double gamma = B.gamma[p];
    B.x[p] += dt * (px / gamma - 1);
    B.y[p] += dt * py / gamma;
    B.z[p] += dt * pz / gamma;
    int emitted = 0;
    if (S.is_photon[s]) {
        B.gamma[p] = sqrt(px * px + py * py + pz * pz);
    } else {
        const double cmr = S.charge_to_mass[s];
        const double pex = B.ex[p], pey = B.ey[p], pez = B.ez[p], pby = B.by[p], pbz = B.bz[p];
        const double fx = cmr * (pex + (py * pbz - pz * pby) / gamma);
        const double fy = cmr * (pey - px * pbz / gamma);
        const double fz = cmr * (pez + px * pby / gamma);
        px += dt * fx;
        py += dt * fy;
        pz += dt * fz;
        gamma = sqrt(1 + px * px + py * py + pz * pz);
        double r, accept;
        if (u_emit) {
            r = u_emit[p];
            accept = u_accept[p];
        } else {
            philox_uniform_pair(seed, step, (unsigned int)p, r, accept);
        }
        const double chi = B.chi[p];
        const double w = w_new(gamma, r, chi, Q);
        const double radiation_probability = w * dt * gamma;
        if (radiation_probability > 0.0) atomicMax(X.max_bits, (unsigned long long)__double_as_longlong(radiation_probability));
        if (accept < radiation_probability) {
            if (chi < 0.13333) {
                const double rm = chi / 0.13333;
                r = r * rm;
            }
            const double qx = r * px, qy = r * py, qz = r * pz;
            X.qx[p] = qx;
            X.qy[p] = qy;
            X.qz[p] = qz;
            emitted = 1;
            px -= qx;
            py -= qx;  // sic: the reference subtracts photon.px from py (src/System_3d.cpp:917)
            pz -= qz;
            gamma = sqrt(1 + px * px + py * py + pz * pz);
        }
        B.px[p] = px;
        B.py[p] = py;
        B.pz[p] = pz;
        B.gamma[p] = gamma;
    }
    X.flag[p] = emitted;
}

// photons in particle order (the order a one-thread reference run appends them in): start = exclusive prefix sum of the flags
__global__ void k_bunch_photon_compact(Bunch B, QedScratch X, const int * __restrict__ start, Bunch photons) {
    QPIC_KERNEL_BEGIN(KID_k_bunch_photon_compact);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B.np) return;
    const int at = start[p];
    if (start[p + 1] == at) return;
    const double qx = X.qx[p], qy = X.qy[p], qz = X.qz[p];
    photons.x[at] = B.x[p];
    photons.y[at] = B.y[p];
    photons.z[at] = B.z[p];
    photons.px[at] = qx;
    photons.py[at] = qy;
    photons.pz[at] = qz;
    photons.n[at] = B.n[p];
    photons.gamma[at] = sqrt(qx * qx + qy * qy + qz * qz);
}

}  // namespace qpic
