/* TEST INFRASTRUCTURE ONLY -- the parity oracle.  Never linked, imported or executed by the product
 * path (quetzalpic_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker.
 *
 * Plain-C, single-threaded restatement of QuetzalPIC's per-xi-slice quasistatic solver,
 * System_3d::solve_wakefield (reference src/System_3d.cpp:238-706) with its deposit / gather primitives
 * (src/array_utils.cpp:50-288) and spectral solve (src/System_3d.cpp:686-706, src/Fourier2d.cpp:32-43).
 *
 * PARITY PINNING: the reference ships no tests or golden vectors.  This oracle is pinned
 * against outputs of the reference itself: the unmodified reference sources are compiled by
 * oracle/Makefile (`make ref`) and tests/golden/make_golden.py records its output for the shipped
 * inputs; tests/test_oracle_vs_golden.py checks this restatement against those fixtures.  The FFT is
 * FFTW in the reference (version unpinned, not installed here); oracle/qo_fft.c restates the DFT from
 * its definition and is pinned against numpy.fft -- "parity unpinned at the FFTW boundary" in that sense.
 */
#ifndef QPIC_ORACLE_H
#define QPIC_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qo_ctx qo_ctx;

typedef struct {
    int nx, ny, nz;
    double lx, ly, lz;       /* box size; steps are re-derived as l/n (src/System_3d.cpp:93-95) */
    int magnetic_field_iterations;
    double psi_threshold;    /* 1e-4  (src/System_3d.h:57) */
    double magnetic_field_D; /* 10.0  (src/System_3d.h:59) */
    double threshold_B;      /* 100   (src/System_3d.cpp:326) */
} qo_params;

/* names follow the reference's HDF5 dataset names (src/System_3d.cpp:49-62) */
typedef enum {
    QO_PSI = 0, QO_PSI_PREV, QO_PSI_MIDDLE, QO_RHO, QO_JX, QO_JY, QO_JZ, QO_JY_NEXT, QO_JZ_NEXT,
    QO_DJY_DXI, QO_DJZ_DXI, QO_RHO_ION, QO_FIELD2D_COUNT
} qo_field2d;

typedef enum {
    QO_ASQR = 0, QO_SUSCEPTIBILITY, QO_RHO_BUNCH, QO_JX_BUNCH, QO_EX, QO_EY, QO_EZ, QO_BY, QO_BZ,
    /* per-slice outputs accumulated the way Output_writer::write_slice fills the Fields file */
    QO_PSI_3D, QO_RHO_3D, QO_JX_3D, QO_JY_3D, QO_JZ_3D, QO_FIELD3D_COUNT
} qo_field3d;

typedef enum { QO_P_Y = 0, QO_P_Z, QO_P_PY, QO_P_PZ, QO_P_N, QO_P_GAMMA, QO_P_PX, QO_P_PY_NEXT, QO_P_PZ_NEXT,
               QO_P_Y_MIDDLE, QO_P_Z_MIDDLE, QO_PARTICLE_FIELD_COUNT } qo_particle_field;

qo_ctx * qo_create(const qo_params * params);
void qo_destroy(qo_ctx * ctx);

/* a_sqr(i,j,k) = amplitude * exp(-sum (x-x0)^2 / w^2), node centred (src/System_3d.cpp:805-816, src/profiles.cpp:4-11) */
void qo_set_laser_gaussian(qo_ctx * ctx, double amplitude, double wx, double wy, double wz, double x0, double y0, double z0);
void qo_set_a_sqr(qo_ctx * ctx, const double * a_sqr);
/* NULL => zeros */
void qo_set_bunch_sources(qo_ctx * ctx, const double * rho_bunch, const double * jx_bunch);

/* explicit particle list (y, z, n); momenta start at 0 (src/System_3d.cpp:736-743) */
int qo_init_wake_particles(qo_ctx * ctx, const double * y, const double * z, const double * n, size_t count);
/* regular sub-cell lattice for a constant profile (src/System_3d.cpp:724-748 with constant2d(rho0)) */
int qo_init_wake_particles_uniform(qo_ctx * ctx, int ppcy, int ppcz, double rho0);

/* whole sweep = qo_begin + qo_step_slice * nx + qo_finish (src/System_3d.cpp:262-626) */
int qo_solve_wakefield(qo_ctx * ctx);
void qo_begin(qo_ctx * ctx);
/* runs the xi-loop body for the current slice (starts at nx-1) and moves to the next one.
 * returns 1 when the sweep is over (slice 0 done or stop condition hit), else 0 */
int qo_step_slice(qo_ctx * ctx);
void qo_finish(qo_ctx * ctx);
int qo_current_slice(const qo_ctx * ctx);
int qo_stopped_at(const qo_ctx * ctx); /* -1, or the slice where |B| > threshold_B / NaN tripped */

size_t qo_particle_count(const qo_ctx * ctx);
double * qo_particles(qo_ctx * ctx, qo_particle_field f); /* SoA view, count doubles */
double * qo_field2d_ptr(qo_ctx * ctx, qo_field2d f);      /* ny*nz, z fastest */
double * qo_field3d_ptr(qo_ctx * ctx, qo_field3d f);      /* nx*ny*nz, z fastest, x slowest */

/* primitives exposed for unit tests (src/array_utils.cpp) */
void qo_deposit(double y, double z, double value, double * array, int ny, int nz, double dy, double dz, double oy, double oz);
void qo_deposit_raw(double y, double z, double value, double * array, int ny, int nz, double dy, double dz);
double qo_gather(double y, double z, const double * array, int ny, int nz, double dy, double dz, double oy, double oz);
double qo_gather_yder(double y, double z, const double * array, int ny, int nz, double dy, double dz);
double qo_gather_zder(double y, double z, const double * array, int ny, int nz, double dy, double dz);
/* (nabla_perp^2 - D) u = src on a periodic ny x nz grid, DC mode forced to 0, result UNNORMALISED
 * (caller divides by ny and nz), in place in `inout` (src/System_3d.cpp:686-706) */
void qo_solve_poisson(qo_ctx * ctx, double * inout, double D);


/* ---- beam ("bunch") side of System_3d::run (src/System_3d.cpp:199-236), oracle/qo_bunch.c.  bunch_particle_3d
 * (src/containers.h:45-61) as caller-owned SoA arrays of `count` entries. */
typedef struct {
    size_t count;
    double *x, *y, *z, *px, *py, *pz, *gamma, *n, *ex, *ey, *ez, *by, *bz, *chi;
    int * species_id;
} qo_bunch;
void qo_deposit3d(double x, double y, double z, double value, double * array, int nx, int ny, int nz, double dx, double dy, double dz,
                  double ox, double oy, double oz);
double qo_gather3d(double x, double y, double z, const double * array, int nx, int ny, int nz, double dx, double dy, double dz,
                   double ox, double oy, double oz);
/* deposit_bunch_densities (src/System_3d.cpp:926-946): zeroes and fills rho_bunch, jx_bunch (nx*ny*nz each) */
void qo_bunch_deposit_densities(const qo_bunch * b, const double * species_charge, int nx, int ny, int nz, double dx, double dy, double dz,
                                double * rho_bunch, double * jx_bunch);
/* deposit_wakefield_to_bunch (src/System_3d.cpp:843-860) */
void qo_bunch_gather_fields(qo_bunch * b, int nx, int ny, int nz, double dx, double dy, double dz, const double * ex, const double * ey,
                            const double * ez, const double * by, const double * bz, double field_schwinger);
/* advance_bunch_particles with qed == false (src/System_3d.cpp:862-886) */
void qo_bunch_advance(qo_bunch * b, const double * species_charge_to_mass, const int * species_is_photon, double dt);
/* W_new (src/monte_carlo.cpp:11-41) */
double qo_w_new(double g, double r, double chi, double omega);
/* advance_bunch_particles with qed == true (src/System_3d.cpp:862-921); the two uniform numbers per particle that the reference
 * takes from rand() are supplied by the caller.  photons: capacity >= b->count; returns the number of emitted photons */
size_t qo_bunch_advance_qed(qo_bunch * b, const double * species_charge_to_mass, const int * species_is_photon, double dt, double frequency,
                            const double * u_emit, const double * u_accept, qo_bunch * photons, double * max_probability);

#ifdef __cplusplus
}
#endif

#endif
