/* qpic_b200 -- C ABI of the B200-native per-xi-slice quasistatic wakefield solver.
 *
 * Drop-in replacement for the body of QuetzalPIC's System_3d::solve_wakefield
 * (reference src/System_3d.cpp:238-627).  The reference has no FFI / plugin boundary: the hot path is a
 * private C++ member (src/System_3d.h:29) operating on members of System_3d (src/System_3d.h:67-92).  This
 * header is the boundary a maintainer would bind instead (INTEGRATION.md shows the System_3d-side patch);
 * each entry point names the reference lines it replaces.
 *
 * Conventions
 *   - plain C, no C++/torch types; all pointers are HOST pointers unless the name starts with d_;
 *   - every array is float64, row-major, z fastest: 2-D (ny,nz) = array2d (src/array2d.h:15-16),
 *     3-D (nx,ny,nz) = array3d (src/array3d.h:13-14);
 *   - functions return QPIC_OK (0) or a negative qpic_status; qpic_last_error() gives the message;
 *   - one context drives one CUDA device; a context is not thread-safe;
 *   - there is NO CPU fallback: qpic_create fails with QPIC_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef QPIC_B200_H
#define QPIC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QPIC_ABI_VERSION 3

typedef enum {
    QPIC_OK = 0,
    QPIC_ERR_INVALID = -1,     /* bad argument (mirrors std::invalid_argument, src/System_3d.cpp:77-85) */
    QPIC_ERR_CUDA = -2,        /* CUDA runtime failure / no device */
    QPIC_ERR_STATE = -3,       /* call sequence error (e.g. step before begin) */
    QPIC_ERR_UNSUPPORTED = -4, /* grid size the batched shared-memory transform cannot handle */
    QPIC_ERR_NOMEM = -5
} qpic_status;

typedef struct qpic_ctx qpic_ctx;

/* Scalars of System_3d that the sweep reads (src/System_3d.h:46-59, src/System_3d.cpp:87-95, :326). */
typedef struct {
    int nx, ny, nz;                /* n = static_cast<int>(l / d), computed by the caller as the reference does */
    double lx, ly, lz;             /* box; steps are re-derived as l / n (src/System_3d.cpp:93-95) */
    int magnetic_field_iterations; /* M >= 0 */
    double psi_threshold;          /* 1e-4 */
    double magnetic_field_D;       /* 10.0 */
    double threshold_B;            /* 100.0 */
    int device;                    /* CUDA device ordinal */
    int keep_slices_3d;            /* != 0: accumulate psi,rho,jx,jy,jz per slice on the device, as Output_writer::write_slice
                                      fills Fields_NNN.h5 (src/System_3d.cpp:663-665); needs 5 more (nx,ny,nz) arrays */
    int use_graph;                 /* != 0 (default): replay one CUDA graph per slice; 0: eager launches (debugging) */
    int stop_check_interval;       /* the host polls the device stop flag every this many slices (<=0: 64) */
    int force_general_deposit;     /* != 0: always use the sort-based deposit (valid for any half-step displacement) instead of
                                      trying the sort-free tile deposit first (valid for displacements <= 1 cell, with rollback) */
    /* slab-decomposed sweep of ONE large transverse grid on several GPUs (dist_world > 1, one context per GPU): the context
     * owns rows [dist_rank * ny / dist_world, (dist_rank + 1) * ny / dist_world) and the particles in them.  Power-of-two ny, nz
     * >= 64 with ny / dist_world a multiple of 16 and >= 32; after qpic_create the ranks exchange their regions
     * (qpic_dist_*) before anything else is called. */
    int dist_rank, dist_world;
    size_t particle_capacity;      /* distributed: room for the particles this rank may hold after migration (0: 1.25 x + 64K) */
    /* Fields_xy output (Output_writer with output_xy, src/Output_writer.cpp:147-187): cuts at z = xy_z0 (< 0: lz / 2, the
     * reference default, src/Config_reader.cpp:218-234).  keep_xy_cuts != 0: the (nx, ny) cuts of the per-slice datasets psi, rho,
     * jx, jy, jz are collected on the device slice by slice, so that Fields_xy needs no (nx, ny, nz) per-slice copies. */
    int keep_xy_cuts;
    double xy_z0;
} qpic_params;

/* fills *p with the reference defaults (psi_threshold 1e-4, D 10, threshold_B 100, device 0, graph on) */
void qpic_default_params(qpic_params * p);

/* 2-D working arrays of System_3d (src/System_3d.h:69-72, 84-91) */
typedef enum {
    QPIC_F2_PSI = 0, QPIC_F2_PSI_PREV, QPIC_F2_PSI_MIDDLE, QPIC_F2_RHO, QPIC_F2_JX, QPIC_F2_JY, QPIC_F2_JZ,
    QPIC_F2_JY_NEXT, QPIC_F2_JZ_NEXT, QPIC_F2_RHO_ION,
    QPIC_F2_BY_CUR, QPIC_F2_BZ_CUR, /* by(i), bz(i) of the slice being solved */
    QPIC_F2_COUNT
} qpic_field2d;

/* 3-D arrays; names = HDF5 dataset names (src/System_3d.cpp:49-62) */
typedef enum {
    QPIC_F3_ASQR = 0, QPIC_F3_SUSCEPTIBILITY, QPIC_F3_RHO_BUNCH, QPIC_F3_JX_BUNCH, QPIC_F3_EX, QPIC_F3_EY, QPIC_F3_EZ,
    QPIC_F3_BY, QPIC_F3_BZ,
    QPIC_F3_PSI, QPIC_F3_RHO, QPIC_F3_JX, QPIC_F3_JY, QPIC_F3_JZ, /* only with keep_slices_3d */
    QPIC_F3_COUNT
} qpic_field3d;

/* wake_particle_2d members that persist between slices (src/containers.h:31-43).  gamma and px are kept for inspection: they are
 * those of the last solved slice once the sweep has reached slice 0; mid-sweep they are refreshed every slice only with the
 * environment variable QPIC_TRACK_GAMMA_PX=1 (two scattered stores per particle and slice otherwise spent for nothing). */
typedef enum { QPIC_P_Y = 0, QPIC_P_Z, QPIC_P_PY, QPIC_P_PZ, QPIC_P_N, QPIC_P_GAMMA, QPIC_P_PX, QPIC_P_COUNT } qpic_particle_field;

/* ---- lifetime.  Replaces the allocation part of System_3d::System_3d (src/System_3d.cpp:155-193). */
int qpic_create(const qpic_params * params, qpic_ctx ** out);
void qpic_destroy(qpic_ctx * ctx);
/* message of the last failure on ctx (ctx == NULL: of the last failed qpic_create in this thread) */
const char * qpic_last_error(const qpic_ctx * ctx);
int qpic_abi_version(void);
/* hash of the sources this binary was compiled from (__graft_entry__.build embeds it); the Python loader refuses a library whose
 * hash differs from the sources next to it, so a stale in-tree binary cannot be used silently */
const char * qpic_source_hash(void);

/* ---- inputs */
/* a_sqr(i,j,k) = amplitude * exp(-sum_c (x_c - r0_c)^2 / w_c^2) at nodes x = i*dx ...; replaces init_a_sqr with
 * gaussian3d (src/System_3d.cpp:805-816, src/profiles.cpp:4-11).  Evaluated on the device. */
int qpic_set_laser_gaussian(qpic_ctx * ctx, double amplitude, double wx, double wy, double wz, double x0, double y0, double z0);
/* any other profile: upload nx*ny*nz values (replaces init_a_sqr for arbitrary std::function profiles) */
int qpic_upload_a_sqr(qpic_ctx * ctx, const double * a_sqr);
/* rho_bunch / jx_bunch as left by System_3d::deposit_bunch_densities (src/System_3d.cpp:926-946); NULL, NULL => zeros */
int qpic_upload_bunch_sources(qpic_ctx * ctx, const double * rho_bunch, const double * jx_bunch);
/* the lattice of System_3d::init_wake_particles (src/System_3d.cpp:724-748), built by the caller: positions and
 * weights n; momenta start at 0.  Particles with n == 0 must not be passed (the reference skips them, :735). */
int qpic_init_wake_particles(qpic_ctx * ctx, const double * y, const double * z, const double * n, size_t count);

/* ---- the sweep.  qpic_solve_wakefield == qpic_begin + qpic_step(nx) + qpic_finish. */
/* whole of solve_wakefield's numerics (src/System_3d.cpp:262-626); returns QPIC_OK also when the |B| stop
 * condition tripped (the reference exits 0 too) -- query qpic_stopped_at. */
int qpic_solve_wakefield(qpic_ctx * ctx);
/* pre-loop zeroing and rho_ion deposition (src/System_3d.cpp:265-326) */
int qpic_begin(qpic_ctx * ctx);
/* up to max_slices iterations of the xi-loop body (src/System_3d.cpp:328-619); asynchronous unless the stop flag
 * has to be polled.  *done (nullable) = 1 once slice 0 has been processed or the stop condition tripped. */
int qpic_step(qpic_ctx * ctx, int max_slices, int * done);
/* output_step(0) (src/System_3d.cpp:622) and a device synchronisation */
int qpic_finish(qpic_ctx * ctx);
/* waits for all queued device work of this context */
int qpic_synchronize(qpic_ctx * ctx);
/* slice index the next qpic_step will process */
int qpic_current_slice(qpic_ctx * ctx);
/* -1, or the slice at which |By|,|Bz| > threshold_B or NaN tripped (src/System_3d.cpp:545-550, 573-578, 588-591) */
int qpic_stopped_at(qpic_ctx * ctx);

/* ---- results / state (each call synchronises) */
int qpic_download_field3d(qpic_ctx * ctx, qpic_field3d f, double * dst);              /* nx*ny*nz doubles */
int qpic_download_slice3d(qpic_ctx * ctx, qpic_field3d f, int slice, double * dst);   /* ny*nz doubles */
int qpic_download_field2d(qpic_ctx * ctx, qpic_field2d f, double * dst);              /* ny*nz doubles */
int qpic_upload_field2d(qpic_ctx * ctx, qpic_field2d f, const double * src);          /* state injection for per-slice parity tests */
size_t qpic_particle_count(const qpic_ctx * ctx);
int qpic_download_particles(qpic_ctx * ctx, qpic_particle_field f, double * dst);     /* count doubles, caller's original order */
int qpic_upload_particles(qpic_ctx * ctx, qpic_particle_field f, const double * src); /* y,z,py,pz,n only */
/* forces the slice index the next qpic_step processes (state injection) */
int qpic_set_current_slice(qpic_ctx * ctx, int slice);

/* ---- beam ("bunch") macroparticles: the part of System_3d::run around solve_wakefield (src/System_3d.cpp:199-236).  The
 * particles (bunch_particle_3d, src/containers.h:45-61) live on the device as SoA arrays; rho_bunch / jx_bunch and the five
 * field arrays they interact with are the resident arrays of the sweep, so a time iteration moves nothing over PCIe.
 * One time iteration of the reference is
 *     qpic_bunch_deposit_densities; qpic_init_wake_particles; qpic_solve_wakefield; qpic_bunch_gather_fields;
 *     [Bunch_NNN output: qpic_bunch_download]; qpic_bunch_advance (all but the last iteration).
 * With qed = true the last call is qpic_bunch_advance_qed, followed by the Photons_NNN output (qpic_photons_download). */
typedef enum {
    QPIC_B_X = 0, QPIC_B_Y, QPIC_B_Z, QPIC_B_PX, QPIC_B_PY, QPIC_B_PZ, QPIC_B_GAMMA, QPIC_B_N,
    QPIC_B_EX, QPIC_B_EY, QPIC_B_EZ, QPIC_B_BY, QPIC_B_BZ, QPIC_B_CHI, QPIC_B_COUNT
} qpic_bunch_field;
/* the Species_register (src/Species_register.h:6-27, built-ins src/Config_reader.cpp:19-22) by species id; at most 16 entries */
int qpic_bunch_set_species(qpic_ctx * ctx, int count, const double * charge, const double * charge_to_mass_ratio, const int * is_photon);
/* the particles System_3d::init_bunch_particles builds (src/System_3d.cpp:767-803), concatenated over the bunches; the field
 * members and chi start at 0.  Must precede the first qpic_step of the context (rho_bunch / jx_bunch become resident). */
int qpic_bunch_init(qpic_ctx * ctx, size_t count, const double * x, const double * y, const double * z, const double * px, const double * py,
                    const double * pz, const double * gamma, const double * n, const int * species_id);
size_t qpic_bunch_count(const qpic_ctx * ctx);
/* System_3d::deposit_bunch_densities (src/System_3d.cpp:926-946) with the 8-point deposit of src/array_utils.cpp:290-338:
 * deterministic (cell lists + one thread per node), no floating-point atomics */
int qpic_bunch_deposit_densities(qpic_ctx * ctx);
/* System_3d::deposit_wakefield_to_bunch (src/System_3d.cpp:843-860): ex, ey, ez, by, bz at the particles (8-point gather
 * honouring each array's origin, src/array_utils.cpp:340-375) and chi; field_schwinger = Plasma_units::field_schwinger */
int qpic_bunch_gather_fields(qpic_ctx * ctx, double field_schwinger);
/* System_3d::advance_bunch_particles with qed == false (src/System_3d.cpp:862-886) */
int qpic_bunch_advance(qpic_ctx * ctx, double dt);
/* System_3d::advance_bunch_particles with qed == true (src/System_3d.cpp:862-921, W_new of src/monte_carlo.cpp:11-41, incl. the
 * `p.py -= photon.px` of :917).  The reference draws two rand_double() per charged particle, in particle order when it runs on
 * one thread: u_emit[p] is the photon energy fraction r (:890), u_accept[p] is compared with the emission probability (:897);
 * host arrays of qpic_bunch_count doubles (entries of photon-species particles are ignored).  Both NULL: a counter-based
 * generator on the device (Philox4x32-10 keyed by `seed`, the step number and the particle index).  frequency =
 * Plasma_units::frequency.  Outputs (nullable): the reference's max_radiation_probability reduction and the number of photons
 * emitted in this step; those photons (in particle order) stay available through qpic_photons_download until the next step. */
int qpic_bunch_advance_qed(qpic_ctx * ctx, double dt, double frequency, const double * u_emit, const double * u_accept, unsigned long long seed,
                           double * max_radiation_probability, size_t * photon_count);
size_t qpic_photon_count(const qpic_ctx * ctx);
/* x, y, z, px, py, pz, gamma, n of the photons of the last step; the field members and chi of a new photon are 0 */
int qpic_photons_download(qpic_ctx * ctx, qpic_bunch_field f, double * dst);
int qpic_bunch_download(qpic_ctx * ctx, qpic_bunch_field f, double * dst); /* qpic_bunch_count doubles */
int qpic_bunch_upload(qpic_ctx * ctx, qpic_bunch_field f, const double * src);

/* ---- building blocks exposed for unit tests and the FFT comparator bench */
/* (nabla_perp^2 - D) u = src on the periodic ny x nz grid with the 5-point-Laplacian symbol, DC forced to 0, result
 * normalised by 1/ny/nz (src/System_3d.cpp:686-706 + the division at :361).  In place on a host array. */
int qpic_solve_poisson_host(qpic_ctx * ctx, double * inout, double D);
/* By/Bz-style pair of Helmholtz solves done as ONE complex transform (re = first field, im = second) */
int qpic_solve_helmholtz_pair_host(qpic_ctx * ctx, double * inout_a, double * inout_b, double D);
/* CUDA-event timing of the last qpic_step call (ms), and kernels launched by it */
int qpic_last_step_stats(qpic_ctx * ctx, float * ms, long * kernel_launches);
/* how many times the sort-free deposit had to be abandoned for a slice (half-step displacement > 1 cell) */
long qpic_fast_path_aborts(const qpic_ctx * ctx);
/* runs up to nslices slices eagerly with a CUDA event after every kernel; report = lines "kernel count total_ms" */
int qpic_profile_slices(qpic_ctx * ctx, int nslices, char * report, size_t report_len);
/* development build only (-DQPIC_TIMELINE): globaltimer stamps left by every kernel of the slices run since the reset;
 * records = 4 values per entry (kernel id, kind 0 = start / 1 = end, t0, t1 in ns); returns the entry count or -1 */
int qpic_timeline_reset(qpic_ctx * ctx);
long qpic_timeline_read(qpic_ctx * ctx, long long * records, long cap);
const char * qpic_timeline_kernel_name(int kernel_id);

/* (count, ny) cut at z = xy_z0 of slices first_slice .. first_slice + count - 1 of a dataset: calculate_xy_slice /
 * calculate_y_slice (src/array_utils.cpp:7-45) evaluated on the device, honouring the array's z-origin (by, ez, jz: dz / 2).
 * Asynchronous like qpic_download_slices3d (copy stream; qpic_wait_transfers before reading dst).  psi, rho, jx, jy, jz need
 * keep_xy_cuts or keep_slices_3d. */
int qpic_download_xy_cut(qpic_ctx * ctx, qpic_field3d f, int first_slice, int count, double * dst);

/* ---- slab-decomposed multi-GPU sweep (qpic_params::dist_world > 1).  Each rank allocates an "exchange region" of
 * identical layout; the ranks map each other's regions and from then on exchange transform transposes, halo rows and
 * migrating particles with peer loads / stores over NVLink inside the slice graph.  Between processes: every rank calls
 * qpic_dist_ipc_handle, the 64-byte handles are all-gathered by the host (MPI / torch.distributed), then every rank calls
 * qpic_dist_attach_ipc with the handles in rank order.  Inside one process: qpic_dist_attach_pointers with the
 * qpic_dist_region pointers of all ranks.  Fields are downloaded per rank (qpic_download_*): own rows are valid.
 * qpic_init_wake_particles takes the particles of the rank's own rows only. */
size_t qpic_dist_region_bytes(const qpic_ctx * ctx);
void * qpic_dist_region(qpic_ctx * ctx);
int qpic_dist_ipc_handle(qpic_ctx * ctx, void * handle64);
int qpic_dist_attach_ipc(qpic_ctx * ctx, const void * handles);
int qpic_dist_attach_pointers(qpic_ctx * ctx, void * const * regions);
/* first own row / number of own rows (whole grid on a single GPU) */
int qpic_dist_row_range(const qpic_ctx * ctx, int * row0, int * nrows);

/* ---- streaming transfers for callers whose arrays live in host memory (System_3d's array3d members).  They run on a
 * separate copy stream and overlap the sweep: an upload is waited for by every slice queued after the call, a download
 * waits for every slice queued before the call.  Call qpic_wait_transfers (or qpic_synchronize) before touching the
 * host buffers.  Pinned buffers (qpic_alloc_pinned) give full PCIe rate and true asynchrony. */
/* a_sqr(first_slice .. first_slice+count-1, :, :)  (the laser envelope is read one slice at a time, src/System_3d.cpp:371) */
int qpic_upload_a_sqr_slices(qpic_ctx * ctx, int first_slice, int count, const double * src);
/* rows [first_slice, first_slice+count) of a 3-D field, e.g. what deposit_wakefield_to_bunch / write_array read afterwards */
int qpic_download_slices3d(qpic_ctx * ctx, qpic_field3d f, int first_slice, int count, double * dst);
int qpic_wait_transfers(qpic_ctx * ctx);
void * qpic_alloc_pinned(size_t bytes);
void qpic_free_pinned(void * p);

#ifdef __cplusplus
}
#endif

#endif /* QPIC_B200_H */
