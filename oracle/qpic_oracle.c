/* TEST INFRASTRUCTURE ONLY -- see qpic_oracle.h for scope and pinning.
 *
 * Single-threaded plain-C restatement of System_3d::solve_wakefield.  Every block cites the lines of
 * /root/reference/src it follows.  Particle loops run in index order, so (unlike the OpenMP reference,
 * whose atomic adds land in arbitrary order) results are bit-reproducible.
 */
#include "qpic_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "qo_fft.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct qo_ctx {
    qo_params p;
    double dx, dy, dz;
    size_t np;
    double * part[QO_PARTICLE_FIELD_COUNT];
    double * f2[QO_FIELD2D_COUNT];
    double * f3[QO_FIELD3D_COUNT];
    double * fin;  /* Fourier2d::in  */
    double * fout; /* Fourier2d::out */
    qo_fft2d fft;
    int cur;        /* current slice index */
    int stop_flag;
    int stopped_at;
    int done;
};

/* ---------------------------------------------------------------- primitives (src/array_utils.cpp) */

/* origin-aware CIC scatter, src/array_utils.cpp:85-118 (2-D) and :50-83 (slice of a 3-D array) */
void qo_deposit(double y, double z, double value, double * a, int ny, int nz, double dy, double dz, double oy, double oz) {
    int j1 = (int)floor((y - oy) / dy);
    int j2 = (j1 + 1) % ny;
    double y_frac = (y - oy) / dy - j1;
    if (j1 < 0) j1 += ny;
    int k1 = (int)floor((z - oz) / dz);
    int k2 = (k1 + 1) % nz;
    double z_frac = (z - oz) / dz - k1;
    if (k1 < 0) k1 += nz;
    a[nz * j1 + k1] += value * (1 - y_frac) * (1 - z_frac);
    a[nz * j2 + k1] += value * y_frac * (1 - z_frac);
    a[nz * j1 + k2] += value * (1 - y_frac) * z_frac;
    a[nz * j2 + k2] += value * y_frac * z_frac;
}

/* raw-pointer overload: no origin, (int) truncation, src/array_utils.cpp:120-142 */
void qo_deposit_raw(double y, double z, double value, double * a, int ny, int nz, double dy, double dz) {
    int j1 = (int)(y / dy);
    int j2 = (j1 + 1) % ny;
    double y_frac = (y / dy) - j1;
    int k1 = (int)(z / dz);
    int k2 = (k1 + 1) % nz;
    double z_frac = (z / dz) - k1;
    a[nz * j1 + k1] += value * (1 - y_frac) * (1 - z_frac);
    a[nz * j2 + k1] += value * y_frac * (1 - z_frac);
    a[nz * j1 + k2] += value * (1 - y_frac) * z_frac;
    a[nz * j2 + k2] += value * y_frac * z_frac;
}

/* origin-aware CIC gather, src/array_utils.cpp:167-188 (2-D) and :144-165 (3-D slice) */
double qo_gather(double y, double z, const double * a, int ny, int nz, double dy, double dz, double oy, double oz) {
    int j1 = (int)floor((y - oy) / dy);
    int j2 = (j1 + 1) % ny;
    double y_frac = (y - oy) / dy - j1;
    if (j1 < 0) j1 += ny;
    int k1 = (int)floor((z - oz) / dz);
    int k2 = (k1 + 1) % nz;
    double z_frac = (z - oz) / dz - k1;
    if (k1 < 0) k1 += nz;
    return a[nz * j1 + k1] * (1 - y_frac) * (1 - z_frac) + a[nz * j2 + k1] * y_frac * (1 - z_frac) +
           a[nz * j1 + k2] * (1 - y_frac) * z_frac + a[nz * j2 + k2] * y_frac * z_frac;
}

/* d/dy at half points, ignores origin, src/array_utils.cpp:240-263 (2-D) and :190-213 (3-D slice) */
double qo_gather_yder(double y, double z, const double * a, int ny, int nz, double dy, double dz) {
    int j1 = (int)floor(y / dy - 0.5);
    double y_frac = (y / dy - 0.5) - j1;
    if (j1 < 0) j1 += ny;
    int j2 = (j1 + 1) % ny;
    int j3 = (j1 + 2) % ny;
    int k1 = (int)(z / dz);
    int k2 = (k1 + 1) % nz;
    double z_frac = (z / dz) - k1;
    double d11 = (a[nz * j2 + k1] - a[nz * j1 + k1]) / dy;
    double d21 = (a[nz * j3 + k1] - a[nz * j2 + k1]) / dy;
    double d12 = (a[nz * j2 + k2] - a[nz * j1 + k2]) / dy;
    double d22 = (a[nz * j3 + k2] - a[nz * j2 + k2]) / dy;
    return d11 * (1 - y_frac) * (1 - z_frac) + d21 * y_frac * (1 - z_frac) + d12 * (1 - y_frac) * z_frac + d22 * y_frac * z_frac;
}

/* d/dz at half points, src/array_utils.cpp:265-288 (2-D) and :215-238 (3-D slice) */
double qo_gather_zder(double y, double z, const double * a, int ny, int nz, double dy, double dz) {
    int j1 = (int)(y / dy);
    int j2 = (j1 + 1) % ny;
    double y_frac = (y / dy) - j1;
    int k1 = (int)floor(z / dz - 0.5);
    double z_frac = (z / dz - 0.5) - k1;
    if (k1 < 0) k1 += nz;
    int k2 = (k1 + 1) % nz;
    int k3 = (k1 + 2) % nz;
    double d11 = (a[nz * j1 + k2] - a[nz * j1 + k1]) / dz;
    double d21 = (a[nz * j2 + k2] - a[nz * j2 + k1]) / dz;
    double d12 = (a[nz * j1 + k3] - a[nz * j1 + k2]) / dz;
    double d22 = (a[nz * j2 + k3] - a[nz * j2 + k2]) / dz;
    return d11 * (1 - y_frac) * (1 - z_frac) + d21 * y_frac * (1 - z_frac) + d12 * (1 - y_frac) * z_frac + d22 * y_frac * z_frac;
}

/* ---------------------------------------------------------------- small helpers (src/System_3d.cpp) */

/* src/System_3d.cpp:30-47: subtract the mean over the first column and the first row */
static void normalize_potential(double * psi, int n1, int n2) {
    double border = 0.0;
    for (int i = 0; i < n1; i++) border += psi[n2 * i];
    for (int j = 1; j < n2; j++) border += psi[j];
    border /= (n1 + n2 - 1);
    for (int i = 0; i < n1 * n2; i++) psi[i] -= border;
}

/* src/System_3d.cpp:830-841 */
static void increase_minimum(double * a, int count, double value) {
    for (int i = 0; i < count; i++) {
        if (a[i] < value) a[i] = value;
    }
}

/* src/System_3d.cpp:668-684 */
static void normalize_coordinates(const qo_ctx * c, double * y, double * z) {
    *y = fmod(*y, c->p.ly);
    if (*y < 0) *y += c->p.ly;
    *z = fmod(*z, c->p.lz);
    if (*z < 0) *z += c->p.lz;
}

/* src/System_3d.cpp:686-706 + src/Fourier2d.cpp:37-43; operates on c->fin */
static void solve_poisson_equation(qo_ctx * c, double D) {
    const int ny = c->p.ny, nz = c->p.nz;
    const int nz_size = nz / 2 + 1;
    qo_rfft2_forward(&c->fft, c->fin, c->fout);
    c->fout[0] = 0;
    c->fout[1] = 0;
    for (int j = 0; j < ny; j++) {
        for (int k = 0; k < nz_size; k++) {
            if (k == 0 && j == 0) continue;
            double multiplier =
                -D + (2 * cos(2 * M_PI * j / ny) - 2) / c->dy / c->dy + (2 * cos(2 * M_PI * k / nz) - 2) / c->dz / c->dz;
            c->fout[2 * (nz_size * j + k)] /= multiplier;
            c->fout[2 * (nz_size * j + k) + 1] /= multiplier;
        }
    }
    qo_rfft2_backward(&c->fft, c->fout, c->fin);
}

void qo_solve_poisson(qo_ctx * c, double * inout, double D) {
    const size_t n = (size_t)c->p.ny * c->p.nz;
    memcpy(c->fin, inout, sizeof(double) * n);
    solve_poisson_equation(c, D);
    memcpy(inout, c->fin, sizeof(double) * n);
}

/* ---------------------------------------------------------------- lifetime / setup */

qo_ctx * qo_create(const qo_params * params) {
    if (!params || params->nx < 1 || params->ny < 1 || params->nz < 1 || params->magnetic_field_iterations < 0) return NULL;
    qo_ctx * c = (qo_ctx *)calloc(1, sizeof(qo_ctx));
    if (!c) return NULL;
    c->p = *params;
    c->dx = params->lx / params->nx;
    c->dy = params->ly / params->ny;
    c->dz = params->lz / params->nz;
    const size_t n2 = (size_t)params->ny * params->nz;
    const size_t n3 = n2 * params->nx;
    for (int f = 0; f < QO_FIELD2D_COUNT; f++) c->f2[f] = (double *)calloc(n2, sizeof(double));
    for (int f = 0; f < QO_FIELD3D_COUNT; f++) c->f3[f] = (double *)calloc(n3, sizeof(double));
    c->fin = (double *)calloc(n2, sizeof(double));
    c->fout = (double *)calloc(2 * (size_t)params->ny * (params->nz / 2 + 1), sizeof(double));
    if (qo_fft2d_init(&c->fft, params->ny, params->nz) != 0) {
        qo_destroy(c);
        return NULL;
    }
    c->cur = params->nx - 1;
    c->stopped_at = -1;
    return c;
}

static void free_particles(qo_ctx * c) {
    for (int f = 0; f < QO_PARTICLE_FIELD_COUNT; f++) {
        free(c->part[f]);
        c->part[f] = NULL;
    }
    c->np = 0;
}

void qo_destroy(qo_ctx * c) {
    if (!c) return;
    free_particles(c);
    for (int f = 0; f < QO_FIELD2D_COUNT; f++) free(c->f2[f]);
    for (int f = 0; f < QO_FIELD3D_COUNT; f++) free(c->f3[f]);
    free(c->fin);
    free(c->fout);
    qo_fft2d_free(&c->fft);
    free(c);
}

/* src/System_3d.cpp:805-816 with src/profiles.cpp:4-11 */
void qo_set_laser_gaussian(qo_ctx * c, double amplitude, double wx, double wy, double wz, double x0, double y0, double z0) {
    double * a = c->f3[QO_ASQR];
    const int nx = c->p.nx, ny = c->p.ny, nz = c->p.nz;
    for (int i = 0; i < nx; i++) {
        for (int j = 0; j < ny; j++) {
            for (int k = 0; k < nz; k++) {
                double x = i * c->dx;
                double y = j * c->dy;
                double z = k * c->dz;
                a[((size_t)i * ny + j) * nz + k] =
                    amplitude * exp(-(x - x0) * (x - x0) / wx / wx - (y - y0) * (y - y0) / wy / wy - (z - z0) * (z - z0) / wz / wz);
            }
        }
    }
}

void qo_set_a_sqr(qo_ctx * c, const double * a_sqr) {
    memcpy(c->f3[QO_ASQR], a_sqr, sizeof(double) * (size_t)c->p.nx * c->p.ny * c->p.nz);
}

void qo_set_bunch_sources(qo_ctx * c, const double * rho_bunch, const double * jx_bunch) {
    const size_t n3 = (size_t)c->p.nx * c->p.ny * c->p.nz;
    if (rho_bunch) memcpy(c->f3[QO_RHO_BUNCH], rho_bunch, sizeof(double) * n3);
    else memset(c->f3[QO_RHO_BUNCH], 0, sizeof(double) * n3);
    if (jx_bunch) memcpy(c->f3[QO_JX_BUNCH], jx_bunch, sizeof(double) * n3);
    else memset(c->f3[QO_JX_BUNCH], 0, sizeof(double) * n3);
}

static int alloc_particles(qo_ctx * c, size_t count) {
    free_particles(c);
    for (int f = 0; f < QO_PARTICLE_FIELD_COUNT; f++) {
        c->part[f] = (double *)calloc(count ? count : 1, sizeof(double));
        if (!c->part[f]) return -1;
    }
    c->np = count;
    /* struct defaults, src/containers.h:31-43: gamma = 1, everything else 0; init sets gamma = 0 (:742) */
    return 0;
}

int qo_init_wake_particles(qo_ctx * c, const double * y, const double * z, const double * n, size_t count) {
    if (alloc_particles(c, count) != 0) return -1;
    memcpy(c->part[QO_P_Y], y, sizeof(double) * count);
    memcpy(c->part[QO_P_Z], z, sizeof(double) * count);
    memcpy(c->part[QO_P_N], n, sizeof(double) * count);
    return 0;
}

/* src/System_3d.cpp:708-748 specialised to constant2d(rho0) (src/profiles.cpp:47-51) */
int qo_init_wake_particles_uniform(qo_ctx * c, int ppcy, int ppcz, double rho0) {
    if (ppcy <= 0 || ppcz <= 0) return -1;
    if (rho0 == 0) return alloc_particles(c, 0);
    const size_t count = (size_t)ppcy * c->p.ny * (size_t)ppcz * c->p.nz;
    if (alloc_particles(c, count) != 0) return -1;
    size_t index = 0;
    for (int i = 0; i < ppcy * c->p.ny; i++) {
        for (int j = 0; j < ppcz * c->p.nz; j++) {
            c->part[QO_P_Y][index] = (i + 0.5) * c->dy / ppcy;
            c->part[QO_P_Z][index] = (j + 0.5) * c->dz / ppcz;
            c->part[QO_P_N][index] = rho0 / ppcy / ppcz;
            index++;
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- the sweep */

/* pre-loop part of solve_wakefield, src/System_3d.cpp:265-326 */
void qo_begin(qo_ctx * c) {
    const int nx = c->p.nx, ny = c->p.ny, nz = c->p.nz;
    const size_t n2 = (size_t)ny * nz;
    const size_t n3 = n2 * nx;
    memset(c->f3[QO_EX] + n2 * (size_t)(nx - 1), 0, sizeof(double) * n2); /* :265-270 */
    memset(c->f2[QO_PSI], 0, sizeof(double) * n2);                        /* :272-277 */
    memset(c->f3[QO_BY], 0, sizeof(double) * n3);                         /* :279-286 */
    memset(c->f3[QO_BZ], 0, sizeof(double) * n3);                         /* :288-295 */
    memset(c->f2[QO_RHO_ION], 0, sizeof(double) * n2);                    /* :297-302 */
    memset(c->f2[QO_JY], 0, sizeof(double) * n2);                         /* :304-309 */
    memset(c->f2[QO_JZ], 0, sizeof(double) * n2);                         /* :311-316 */
    /* momenta restart from zero because init_wake_particles re-runs every sweep (:262, :736-743) */
    for (size_t pi = 0; pi < c->np; pi++) {                               /* :319-323 */
        qo_deposit(c->part[QO_P_Y][pi], c->part[QO_P_Z][pi], -c->part[QO_P_N][pi], c->f2[QO_RHO_ION], ny, nz, c->dy, c->dz, 0.0, 0.0);
    }
    c->stop_flag = 0;
    c->stopped_at = -1;
    c->cur = nx - 1;
    c->done = 0;
}

/* src/System_3d.cpp:629-666; write_slice == store the five 2-D arrays into row i of the 3-D outputs */
static void output_step(qo_ctx * c, int i) {
    const int nx = c->p.nx, ny = c->p.ny, nz = c->p.nz;
    const size_t n2 = (size_t)ny * nz;
    const double * psi = c->f2[QO_PSI];
    const double * psi_prev = c->f2[QO_PSI_PREV];
    double * ex = c->f3[QO_EX] + n2 * i;
    double * ey = c->f3[QO_EY] + n2 * i;
    double * ez = c->f3[QO_EZ] + n2 * i;
    const double * by = c->f3[QO_BY] + n2 * i;
    const double * bz = c->f3[QO_BZ] + n2 * i;
    if (i < nx - 1) {
        for (size_t t = 0; t < n2; t++) ex[t] = (psi[t] - psi_prev[t]) / c->dx;
    }
    for (int j = 0; j < ny - 1; j++) {
        for (int k = 0; k < nz; k++) ey[nz * j + k] = -(psi[nz * (j + 1) + k] - psi[nz * j + k]) / c->dy + bz[nz * j + k];
    }
    for (int k = 0; k < nz; k++) ey[nz * (ny - 1) + k] = -(psi[k] - psi[nz * (ny - 1) + k]) / c->dy + bz[nz * (ny - 1) + k];
    for (int j = 0; j < ny; j++) {
        for (int k = 0; k < nz - 1; k++) ez[nz * j + k] = -(psi[nz * j + k + 1] - psi[nz * j + k]) / c->dz - by[nz * j + k];
        ez[nz * j + nz - 1] = -(psi[nz * j] - psi[nz * j + nz - 1]) / c->dz - by[nz * j + nz - 1];
    }
    memcpy(c->f3[QO_PSI_3D] + n2 * i, c->f2[QO_PSI], sizeof(double) * n2);
    memcpy(c->f3[QO_RHO_3D] + n2 * i, c->f2[QO_RHO], sizeof(double) * n2);
    memcpy(c->f3[QO_JX_3D] + n2 * i, c->f2[QO_JX], sizeof(double) * n2);
    memcpy(c->f3[QO_JY_3D] + n2 * i, c->f2[QO_JY], sizeof(double) * n2);
    memcpy(c->f3[QO_JZ_3D] + n2 * i, c->f2[QO_JZ], sizeof(double) * n2);
}

/* fourier.in <- jx_bunch(i) - rho_bunch(i) - rho_ion, then the raw deposit of -n; :342-353 and :446-457 */
static void build_psi_source(qo_ctx * c, int i, const double * py_, const double * pz_) {
    const int ny = c->p.ny, nz = c->p.nz;
    const size_t n2 = (size_t)ny * nz;
    const double * jxb = c->f3[QO_JX_BUNCH] + n2 * i;
    const double * rhob = c->f3[QO_RHO_BUNCH] + n2 * i;
    const double * rho_ion = c->f2[QO_RHO_ION];
    for (size_t t = 0; t < n2; t++) c->fin[t] = jxb[t] - rhob[t] - rho_ion[t];
    for (size_t pi = 0; pi < c->np; pi++) qo_deposit_raw(py_[pi], pz_[pi], -c->part[QO_P_N][pi], c->fin, ny, nz, c->dy, c->dz);
}

/* :358-365 / :462-469 */
static void extract_potential(qo_ctx * c, double * dst) {
    const int ny = c->p.ny, nz = c->p.nz;
    const size_t n2 = (size_t)ny * nz;
    for (size_t t = 0; t < n2; t++) dst[t] = c->fin[t] / ny / nz;
    normalize_potential(dst, ny, nz);
    increase_minimum(dst, ny * nz, c->p.psi_threshold - 1);
}

/* rho, jx, susceptibility(i) reset + deposit; :381-397 and :511-527 */
static void deposit_rho_jx(qo_ctx * c, int i) {
    const int ny = c->p.ny, nz = c->p.nz;
    const size_t n2 = (size_t)ny * nz;
    double * rho = c->f2[QO_RHO];
    double * jx = c->f2[QO_JX];
    double * chi = c->f3[QO_SUSCEPTIBILITY] + n2 * i;
    const double * jxb = c->f3[QO_JX_BUNCH] + n2 * i;
    const double * rhob = c->f3[QO_RHO_BUNCH] + n2 * i;
    for (size_t t = 0; t < n2; t++) {
        rho[t] = rhob[t];
        jx[t] = jxb[t];
        chi[t] = 0.0;
    }
    const double *y = c->part[QO_P_Y], *z = c->part[QO_P_Z], *n = c->part[QO_P_N], *gamma = c->part[QO_P_GAMMA], *px = c->part[QO_P_PX];
    for (size_t pi = 0; pi < c->np; pi++) {
        double vx = px[pi] / gamma[pi];
        qo_deposit(y[pi], z[pi], n[pi] * vx / (1 - vx), jx, ny, nz, c->dy, c->dz, 0.0, 0.0);
        qo_deposit(y[pi], z[pi], n[pi] / (1 - vx), rho, ny, nz, c->dy, c->dz, 0.0, 0.0);
        qo_deposit(y[pi], z[pi], -n[pi] / (1 - vx) / gamma[pi], chi, ny, nz, c->dy, c->dz, 0.0, 0.0);
    }
}

int qo_step_slice(qo_ctx * c) {
    if (c->done) return 1;
    const int nx = c->p.nx, ny = c->p.ny, nz = c->p.nz;
    const size_t n2 = (size_t)ny * nz;
    const int i = c->cur;
    const double dx = c->dx, dy = c->dy, dz = c->dz;
    const double D = c->p.magnetic_field_D;
    double *y = c->part[QO_P_Y], *z = c->part[QO_P_Z], *py = c->part[QO_P_PY], *pz = c->part[QO_P_PZ];
    double *pyn = c->part[QO_P_PY_NEXT], *pzn = c->part[QO_P_PZ_NEXT], *ym = c->part[QO_P_Y_MIDDLE], *zm = c->part[QO_P_Z_MIDDLE];
    double *n = c->part[QO_P_N], *gamma = c->part[QO_P_GAMMA], *px = c->part[QO_P_PX];
    double *psi = c->f2[QO_PSI], *psi_prev = c->f2[QO_PSI_PREV], *psi_middle = c->f2[QO_PSI_MIDDLE];
    double *jx = c->f2[QO_JX], *jy = c->f2[QO_JY], *jz = c->f2[QO_JZ], *jyn = c->f2[QO_JY_NEXT], *jzn = c->f2[QO_JZ_NEXT];
    double *djy = c->f2[QO_DJY_DXI], *djz = c->f2[QO_DJZ_DXI];
    const double * a_sqr = c->f3[QO_ASQR] + n2 * i;
    double * by = c->f3[QO_BY] + n2 * i;
    double * bz = c->f3[QO_BZ] + n2 * i;

    memcpy(psi_prev, psi, sizeof(double) * n2); /* :334-339 */

    build_psi_source(c, i, y, z); /* :342-353 */
    solve_poisson_equation(c, 0.0); /* :356 */
    extract_potential(c, psi);      /* :358-365 */

    for (size_t pi = 0; pi < c->np; pi++) { /* :368-378 */
        double a_particle = qo_gather(y[pi], z[pi], a_sqr, ny, nz, dy, dz, 0.0, 0.0);
        double psi_particle = qo_gather(y[pi], z[pi], psi, ny, nz, dy, dz, 0.0, 0.0);
        gamma[pi] = 0.5 * (1 + py[pi] * py[pi] + pz[pi] * pz[pi] + a_particle + (1 + psi_particle) * (1 + psi_particle)) / (1 + psi_particle);
        px[pi] = 0.5 * (1 + py[pi] * py[pi] + pz[pi] * pz[pi] + a_particle - (1 + psi_particle) * (1 + psi_particle)) / (1 + psi_particle);
    }

    deposit_rho_jx(c, i); /* :381-397 */

    if (i < nx - 1) { /* :399-407 */
        memcpy(by, by + n2, sizeof(double) * n2);
        memcpy(bz, bz + n2, sizeof(double) * n2);
    }

    if (i == 0) { /* :409-411 */
        c->done = 1;
        return 1;
    }

    for (int iteration = 0; iteration < c->p.magnetic_field_iterations; iteration++) { /* :413 */
        for (size_t pi = 0; pi < c->np; pi++) { /* :416-433 */
            double psi_particle = qo_gather(y[pi], z[pi], psi, ny, nz, dy, dz, 0.0, 0.0);
            double da_dy_particle = qo_gather_yder(y[pi], z[pi], a_sqr, ny, nz, dy, dz);
            double da_dz_particle = qo_gather_zder(y[pi], z[pi], a_sqr, ny, nz, dy, dz);
            double dpsi_dy_particle = qo_gather_yder(y[pi], z[pi], psi, ny, nz, dy, dz);
            double dpsi_dz_particle = qo_gather_zder(y[pi], z[pi], psi, ny, nz, dy, dz);
            double by_particle = qo_gather(y[pi], z[pi], by, ny, nz, dy, dz, 0.0, 0.5 * dz); /* by origin (0,0,dz/2), :180 */
            double bz_particle = qo_gather(y[pi], z[pi], bz, ny, nz, dy, dz, 0.5 * dy, 0.0); /* bz origin (0,dy/2,0), :181 */
            pyn[pi] = py[pi] - dx * 0.5 * da_dy_particle / (1 + psi_particle);
            pyn[pi] += dx * gamma[pi] * dpsi_dy_particle / (1 + psi_particle);
            pyn[pi] -= dx * bz_particle;
            pzn[pi] = pz[pi] - dx * 0.5 * da_dz_particle / (1 + psi_particle);
            pzn[pi] += dx * gamma[pi] * dpsi_dz_particle / (1 + psi_particle);
            pzn[pi] += dx * by_particle;
        }
        for (size_t pi = 0; pi < c->np; pi++) { /* :436-443 */
            double psi_particle = qo_gather(y[pi], z[pi], psi, ny, nz, dy, dz, 0.0, 0.0);
            ym[pi] = y[pi] + 0.5 * dx * pyn[pi] / (1 + psi_particle);
            zm[pi] = z[pi] + 0.5 * dx * pzn[pi] / (1 + psi_particle);
            normalize_coordinates(c, &ym[pi], &zm[pi]);
        }

        build_psi_source(c, i, ym, zm); /* :446-457 */
        solve_poisson_equation(c, 0.0);  /* :460 */
        extract_potential(c, psi_middle); /* :462-469 */

        memset(jyn, 0, sizeof(double) * n2); /* :472-478 */
        memset(jzn, 0, sizeof(double) * n2);
        for (size_t pi = 0; pi < c->np; pi++) { /* :480-486; jy_next origin (dy/2, 0), jz_next (0, dz/2), :190-193 */
            double psi_particle = qo_gather(ym[pi], zm[pi], psi_middle, ny, nz, dy, dz, 0.0, 0.0);
            qo_deposit(ym[pi], zm[pi], n[pi] * pyn[pi] / (1 + psi_particle), jyn, ny, nz, dy, dz, 0.5 * dy, 0.0);
            qo_deposit(ym[pi], zm[pi], n[pi] * pzn[pi] / (1 + psi_particle), jzn, ny, nz, dy, dz, 0.0, 0.5 * dz);
        }
        for (size_t t = 0; t < n2; t++) { /* :489-495 */
            djy[t] = (jy[t] - jyn[t]) / dx;
            djz[t] = (jz[t] - jzn[t]) / dx;
        }
        for (size_t pi = 0; pi < c->np; pi++) { /* :498-508 */
            double a_particle = qo_gather(y[pi], z[pi], a_sqr, ny, nz, dy, dz, 0.0, 0.0);
            double psi_particle = qo_gather(y[pi], z[pi], psi, ny, nz, dy, dz, 0.0, 0.0);
            double p_squared = 0.25 * (py[pi] + pyn[pi]) * (py[pi] + pyn[pi]) + 0.25 * (pz[pi] + pzn[pi]) * (pz[pi] + pzn[pi]);
            gamma[pi] = 0.5 * (1 + p_squared + a_particle + (1 + psi_particle) * (1 + psi_particle)) / (1 + psi_particle);
            px[pi] = 0.5 * (1 + p_squared + a_particle - (1 + psi_particle) * (1 + psi_particle)) / (1 + psi_particle);
        }

        deposit_rho_jx(c, i); /* :511-527 */

        for (int j = 0; j < ny; j++) { /* :530-536 */
            for (int k = 0; k < nz - 1; k++) {
                c->fin[nz * j + k] = -D * by[nz * j + k] - (jx[nz * j + k + 1] - jx[nz * j + k]) / dz + djz[nz * j + k];
            }
            c->fin[nz * j + (nz - 1)] = -D * by[nz * j + nz - 1] - (jx[nz * j] - jx[nz * j + nz - 1]) / dz + djz[nz * j + nz - 1];
        }
        solve_poisson_equation(c, D); /* :539 */
        for (size_t t = 0; t < n2; t++) { /* :541-552 */
            by[t] = c->fin[t] / ny / nz;
            if (fabs(by[t]) > c->p.threshold_B) c->stop_flag = 1;
            if (isnan(by[t])) c->stop_flag = 1;
        }

        for (int j = 0; j < ny - 1; j++) { /* :555-560 */
            for (int k = 0; k < nz; k++) {
                c->fin[nz * j + k] = -D * bz[nz * j + k] + (jx[nz * (j + 1) + k] - jx[nz * j + k]) / dy - djy[nz * j + k];
            }
        }
        for (int k = 0; k < nz; k++) { /* :561-564 -- NB the reference reads bz(i, 0, k) here, reproduced on purpose */
            c->fin[nz * (ny - 1) + k] = -D * bz[k] + (jx[k] - jx[nz * (ny - 1) + k]) / dy - djy[nz * (ny - 1) + k];
        }
        solve_poisson_equation(c, D); /* :567 */
        for (size_t t = 0; t < n2; t++) { /* :569-580 */
            bz[t] = c->fin[t] / ny / nz;
            if (fabs(bz[t]) > c->p.threshold_B) c->stop_flag = 1;
            if (isnan(bz[t])) c->stop_flag = 1;
        }

        if (c->stop_flag) break; /* :582-584 */
    }

    if (c->stop_flag) { /* :588-591 */
        c->stopped_at = i;
        c->done = 1;
        return 1;
    }

    for (size_t pi = 0; pi < c->np; pi++) { /* :593-598 */
        py[pi] = pyn[pi];
        pz[pi] = pzn[pi];
    }
    for (size_t pi = 0; pi < c->np; pi++) { /* :601-608 */
        double psi_particle = qo_gather(y[pi], z[pi], psi_middle, ny, nz, dy, dz, 0.0, 0.0);
        y[pi] += dx * py[pi] / (1 + psi_particle);
        z[pi] += dx * pz[pi] / (1 + psi_particle);
        normalize_coordinates(c, &y[pi], &z[pi]);
    }

    output_step(c, i); /* :610 */

    memcpy(jy, jyn, sizeof(double) * n2); /* :612-618 */
    memcpy(jz, jzn, sizeof(double) * n2);

    c->cur = i - 1;
    return 0;
}

void qo_finish(qo_ctx * c) { output_step(c, 0); /* :622 */ }

int qo_solve_wakefield(qo_ctx * c) {
    qo_begin(c);
    while (!qo_step_slice(c)) {
    }
    qo_finish(c);
    return c->stopped_at;
}

int qo_current_slice(const qo_ctx * c) { return c->cur; }
int qo_stopped_at(const qo_ctx * c) { return c->stopped_at; }
size_t qo_particle_count(const qo_ctx * c) { return c->np; }
double * qo_particles(qo_ctx * c, qo_particle_field f) { return c->part[f]; }
double * qo_field2d_ptr(qo_ctx * c, qo_field2d f) { return c->f2[f]; }
double * qo_field3d_ptr(qo_ctx * c, qo_field3d f) { return c->f3[f]; }
