// qpic_b200: context, kernel sequencing and the C ABI (include/qpic_b200.h).
//
// One context = one CUDA device, one stream.  The whole xi-loop body of System_3d::solve_wakefield
// (reference src/System_3d.cpp:328-619) is a fixed sequence of kernels that read the slice index from device
// memory, so it is captured ONCE into a CUDA graph and replayed nx-1 times; the host only polls the stop flag
// every few dozen slices.  All field and particle arrays stay resident in HBM for the whole sweep.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

#include "../../include/qpic_b200.h"
#include "qpic_compat.h"
#include "qpic_fft.cuh"
#include "qpic_fft_pow2.cuh"
#include "qpic_kernels.cuh"
#include "qpic_dist.cuh"
#include "qpic_bunch.cuh"

// build (MODE) of a power-of-two transform kernel: P2_PICK(mode, kernel, other template arguments...) / P2_PICK0(mode, kernel)
#define P2_PICK(mode, K, ...) ((mode) == 2 ? K<__VA_ARGS__, 2> : ((mode) == 1 ? K<__VA_ARGS__, 1> : K<__VA_ARGS__, 0>))
#define P2_PICK0(mode, K) ((mode) == 2 ? K<2> : ((mode) == 1 ? K<1> : K<0>))

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

using namespace qpic;

#ifndef QPIC_PBLOCK
#define QPIC_PBLOCK 128  // threads per block of the particle kernels (one particle per thread)
#endif

namespace {

thread_local std::string g_create_error;

constexpr size_t SMEM_OPTIN_LIMIT = 227 * 1024;
constexpr size_t SMEM_DEFAULT_LIMIT = 48 * 1024;

struct DeviceFft {
    FftPlan plan;
    double2 * tw = nullptr;  // device
};

}  // namespace

struct qpic_ctx {
    qpic_params prm;
    Geom g;
    std::string err;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // second branch of the per-iteration fork (see enqueue_slice_full)
    cudaStream_t copy_stream = nullptr;  // streaming transfers (a_sqr slabs in, field slabs out) overlap the sweep
    cudaEvent_t ev_copy = nullptr;
    std::vector<cudaEvent_t> ev_pool;
    int ev_next = 0;
    bool fast_capable = false;  // sort-free tile deposits usable for this grid
    int general_left = 0;       // slices still to run on the general path after a fast-path abort
    int general_penalty = 16;
    long fast_aborts = 0;
    int target = 0;             // slice index the caller has asked the sweep to reach (-1: slice 0 processed too)
    bool began = false;
    bool done = false;
    bool finished = false;
    bool lists_dirty = false;
    int cur = 0;          // host mirror of Ctl::slice
    int stopped_at = -1;
    int since_poll = 0;
    // lagged poll (qpic_step without `done`): the control block is copied asynchronously behind the slices queued so far and looked
    // at one polling interval later, so the host never waits for the device while it is only queueing work
    bool lagged = false, snap_pending = false;
    Ctl * ctl_snap = nullptr;  // pinned
    cudaEvent_t ev_snap = nullptr, ev_poll_pos = nullptr;
    cudaStream_t poll_stream = nullptr;  // carries the lagged control-block snapshots (poll_ctl_lagged)

    // device state
    Ctl * ctl = nullptr;
    Ctl * ctl_scratch = nullptr;  // for the exposed building blocks (never carries a stop tag)
    Ctl * ctl_host = nullptr;     // pinned
    double * f2[QPIC_F2_COUNT] = {};
    double * f3[QPIC_F3_COUNT] = {};
    double *fft_in = nullptr, *fft_in2 = nullptr, *raw = nullptr;
    double2 * spec = nullptr;  // ny * nz complex (the psi solves use the first ny*(nz/2+1))
    double2 *spec_row0 = nullptr, *spec_colsum = nullptr;  // nz/2+1 each: U(0,k) and sum_j U(j,k) for the fused potential fix-up
    int * eos_ticket = nullptr;                            // "last block" counter of k_end_of_slice (advances Ctl::slice)
    int * cols_ticket = nullptr;                           // "last block" counter of the column kernel
    double *sy = nullptr, *sz = nullptr;
    DeviceFft fy, fz;
    Particles P = {};
    int *count = nullptr, *cursor = nullptr, *sums = nullptr, *tmp = nullptr;
    int *start_b = nullptr, *start_m = nullptr, *perm_b = nullptr, *perm_m = nullptr;
    int nchunks = 0;
    std::vector<void *> allocations;

    // launch geometry of the transforms: sequences per block, threads, dynamic shared memory
    struct FftLaunch {
        int spb = 1, threads = 128, blocks = 1;
        size_t smem = 0;
        int mode = 0;  // power-of-two engine: build of the kernels (0 narrow, 1 wide, 2 deep: see qpic_fft_pow2.cuh)
    };
    FftLaunch rows_real, rows_cplx, cols_half, cols_full;
    bool pow2_y = false, pow2_z = false;
    double inv_scale = 0.0;  // 1/(ny*nz) when that is exactly representable (power of two), else 0 = divide
    struct InvSymbol {
        double D;
        bool full;
        double * table;  // device, [ncols][ny]
    };
    std::vector<InvSymbol> inv_symbols;  // for the power-of-two column kernel

    // graph
    cudaGraph_t graph = nullptr, graph_fast = nullptr;
    cudaGraphExec_t graph_exec = nullptr, graph_exec_fast = nullptr;
    long launches_per_slice = 0, launches_per_slice_fast = 0;

    // stats
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    long last_launches = 0;
    long launch_counter = 0;

    // slab-decomposed multi-GPU sweep (qpic_dist.cuh)
    struct Dist {
        bool on = false, attached = false;
        DistView v = {};
        char * region = nullptr;
        int cpr_half = 0, cpr_full = 0;            // columns per rank of the half / full spectrum
        int col0_half = 0, ncols_half = 0, col0_full = 0, ncols_full = 0;
        int * out_count = nullptr;                 // emigrant counters [2]
        long long * epoch = nullptr;
        int live0 = 0;                             // particles handed to qpic_init_wake_particles
        std::vector<void *> opened;                // cudaIpcOpenMemHandle mappings
    } dist;
    double2 * b_src = nullptr;                     // sources of the By + i Bz solve formed by the jy/jz deposit (TileJob::b_src)
    int *bin_cnt = nullptr, *bin_lst = nullptr;  // per-cell lists of the half-step positions (TileJob::bin_cnt)
    bool jyjz_lists = true;                        // mover lists of the half-step positions exist (false: force_general_deposit, no fast path)
    // half steps beyond one cell stay on the sort-free path (FarMovers, qpic_kernels.cuh); QPIC_FAR_MOVERS=0: they raise the rollback as in round 1
    bool far_movers = true;
    int * far_cnt[2] = {nullptr, nullptr};
    int * far_lst = nullptr;
    // stayers of the psi_middle source summed by the predictor (k_particles_predict_dep + k_tile_assemble); chosen by problem size, see alloc_particles
    bool predict_deposit = false;
    int predict_deposit_mode = -1;  // -1 auto, 0 off, 1 on
    StayOut stay = {};
    int rc_folded_now = 0;  // set around the slice-end launch of the rho / susceptibility tile kernel, see TileJob::rc_folded
    int * stay_cnt[2] = {nullptr, nullptr};
    int cur_it = 0;
    int fold_rc_mode = -1;                         // ... and rho / susceptibility in the last pass of a slice: QPIC_FOLD_RHO_CHI = 0 / 1, unset: by density
    bool fold_rc = false;
    bool dense_cells = false;                      // >= 8 particles per cell: k_particles_predict_dep<., ., SRED = true>
    bool fold_jx = true;                           // predictor-side deposit: the jx deposit at (y, z) rides along (QPIC_FOLD_JX=0: separate tile kernel)
    bool fuse_gamma = true;                        // the first predictor pass of a slice does k_particles_gamma's work (QPIC_FUSE_GAMMA=0: separate kernel)
    bool track_gamma_px = false;                   // QPIC_TRACK_GAMMA_PX=1: gamma, px by particle id after EVERY slice (default: at the end of the sweep)
    double * xy[5] = {};          // keep_xy_cuts: (nx, ny) cuts of psi, rho, jx, jy, jz
    double * xy_scratch = nullptr;  // qpic_download_xy_cut: (nx, ny) staging for the cuts of resident 3-D arrays
    double * partials = nullptr;  // border-sum partials of the fused potential fix-up, [world][4] (inside the region when distributed)

    // beam macroparticles and the cell lists of their deposit (qpic_bunch.cuh)
    struct BunchState {
        Bunch b = {};
        BunchSpecies species = {};
        int nspecies = 0;
        int capacity = 0;
        int *cell = nullptr, *tmp = nullptr, *perm = nullptr;     // per particle
        int *count = nullptr, *start = nullptr, *sums = nullptr;  // per cell of the (nx, ny, nz) grid / per scan chunk
        int nchunks = 0;
        // photon emission (qed): photons of the last step, scratch of the emission kernel
        Bunch photons = {};
        QedScratch qed = {};
        int *qed_start = nullptr, *qed_sums = nullptr;
        int qed_capacity = 0;
        unsigned int qed_step = 0;
        double *u_emit = nullptr, *u_accept = nullptr;
    } bunch;

    // per-kernel profiling (qpic_profile_slices): an event after every launch of an eager slice
    bool profiling = false;
    std::vector<const char *> prof_names;
    std::vector<cudaEvent_t> prof_events;
};

namespace {

int fail(qpic_ctx * c, int code, const char * fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}

#define QPIC_CUDA(c, call)                                                                                            \
    do {                                                                                                              \
        cudaError_t e_ = (call);                                                                                      \
        if (e_ != cudaSuccess) return fail((c), QPIC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define QPIC_TRY(expr)            \
    do {                          \
        int rc_ = (expr);         \
        if (rc_ != QPIC_OK) return rc_; \
    } while (0)

template <class T>
int dev_alloc(qpic_ctx * c, T ** p, size_t count, bool zero = true) {
    void * raw = nullptr;
    const size_t bytes = (count ? count : 1) * sizeof(T);
    if (*p) {
        // re-allocation (qpic_init_wake_particles with another particle count on a context without a captured graph): the previous buffer
        // is released instead of staying with the context until qpic_destroy
        auto it = std::find(c->allocations.begin(), c->allocations.end(), static_cast<void *>(*p));
        if (it != c->allocations.end()) {
            cudaStreamSynchronize(c->stream);
            cudaFree(*it);
            c->allocations.erase(it);
        }
        *p = nullptr;
    }
    cudaError_t e = cudaMalloc(&raw, bytes);
    if (e != cudaSuccess) return fail(c, QPIC_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    c->allocations.push_back(raw);
    *p = static_cast<T *>(raw);
    if (zero) QPIC_CUDA(c, cudaMemsetAsync(raw, 0, bytes, c->stream));
    return QPIC_OK;
}

int blocks_for(size_t n, int threads) { return (int)((n + threads - 1) / threads); }

int pick_threads(int work) {
    int t = ((work + 31) / 32) * 32;
    if (t < 64) t = 64;
    if (t > 512) t = 512;
    return t;
}

// factorisation 4,4,...,2,odd primes ascending; twiddle table exp(-2 pi i t / n) in extended precision
int make_fft(qpic_ctx * c, int n, DeviceFft * out) {
    FftPlan & P = out->plan;
    P.n = n;
    P.nfac = 0;
    P.log2n = 0;
    if (n >= 64 && (n & (n - 1)) == 0) {
        while ((1 << P.log2n) < n) P.log2n++;
    }
    int m = n;
    auto push = [&](int f) -> bool {
        if (P.nfac >= FFT_MAX_FACTORS) return false;
        P.fac[P.nfac++] = f;
        return true;
    };
    while (m % 4 == 0) {
        if (!push(4)) return fail(c, QPIC_ERR_UNSUPPORTED, "grid size %d has too many factors", n);
        m /= 4;
    }
    while (m % 2 == 0) {
        if (!push(2)) return fail(c, QPIC_ERR_UNSUPPORTED, "grid size %d has too many factors", n);
        m /= 2;
    }
    for (int f = 3; m > 1; f += 2) {
        while (m % f == 0) {
            if (f > 127) return fail(c, QPIC_ERR_UNSUPPORTED, "grid size %d has a prime factor %d > 127; not supported by the shared-memory transform", n, f);
            if (!push(f)) return fail(c, QPIC_ERR_UNSUPPORTED, "grid size %d has too many factors", n);
            m /= f;
        }
    }
    std::vector<double2> tw(n);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int t = 0; t < n; t++) {
        const long double a = two_pi * (long double)t / (long double)n;
        tw[t].x = (double)cosl(a);
        tw[t].y = (double)(-sinl(a));
    }
    QPIC_TRY(dev_alloc(c, &out->tw, (size_t)n, false));
    QPIC_CUDA(c, cudaMemcpyAsync(out->tw, tw.data(), sizeof(double2) * n, cudaMemcpyHostToDevice, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));  // tw is a local
    return QPIC_OK;
}

// nseq sequences of length n: how many per block.  Power-of-two engine: exactly n/8 threads per sequence, in place;
// generic engine: one block-wide loop, ping-pong buffers.
// what a set of kernels that share a launch shape allows per block: threads (register use) and dynamic shared memory (227 KB
// minus their static shared memory)
struct KernelLimits {
    int max_threads = 1024;
    size_t max_dyn_smem = SMEM_OPTIN_LIMIT;
};

template <class... K>
KernelLimits kernel_limits(K... kernels) {
    KernelLimits lim;
#ifndef QPIC_EMU
    const void * list[] = {reinterpret_cast<const void *>(kernels)...};
    for (const void * k : list) {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, k) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        if (fa.maxThreadsPerBlock < lim.max_threads) lim.max_threads = fa.maxThreadsPerBlock;
        const size_t room = SMEM_OPTIN_LIMIT > fa.sharedSizeBytes ? SMEM_OPTIN_LIMIT - fa.sharedSizeBytes : 0;
        if (room < lim.max_dyn_smem) lim.max_dyn_smem = room;
    }
#endif
    return lim;
}

// lim[0]: the narrow builds of the kernels (power-of-two engine) / the generic engine's kernels; lim[1], lim[2]: the wide and deep builds
int size_fft_launch(qpic_ctx * c, const FftPlan & P, int nseq, qpic_ctx::FftLaunch * L, const char * what, const KernelLimits * lim3) {
    const KernelLimits none;
    const KernelLimits & lim = lim3 ? lim3[0] : none;
    const int n = P.n;
    if (P.log2n > 0) {
        const int e8 = n / 8;
        // These kernels are latency bound (a few warps per SM at most), so the fastest shape is ONE wave of fat blocks: as many sequences per
        // block as it takes to cover the batch with <= 148 blocks, within the kernels' thread (register) and shared-memory limits.  Where two
        // alternating exchange buffers per sequence do not leave room for that, a single-buffer build -- one more barrier per exchange -- is
        // used if it saves waves: wide (1024-point axes: 7 sequences per block instead of 4) or deep (4096-point axes: a sequence fills a
        // block, two blocks per SM instead of one).
        const size_t one = sizeof(double2) * (size_t)p2_pad_len(n);
        int best = -1, best_spb = 1;
        long best_waves = 0;
        const char * e_deep = getenv("QPIC_FFT_DEEP");  // A/B: 0 = narrow / wide builds only
        const int modes = lim3 ? ((e_deep && e_deep[0] == '0') ? 2 : 3) : 1;
        for (int mode = 0; mode < modes; mode++) {
            const KernelLimits & lm = lim3 ? lim3[mode] : none;
            const size_t per_seq = mode == 0 ? 2 * one : one;
            int spb = (nseq + 147) / 148;
            while (spb > 1 && (spb * e8 > lm.max_threads || spb * per_seq > lm.max_dyn_smem)) spb--;
            if (spb < 1) spb = 1;
            while (spb * e8 < 32) spb *= 2;
            if (spb * e8 > lm.max_threads || spb * per_seq > lm.max_dyn_smem) continue;
            // resident blocks per SM: shared memory (227 KB per SM, 1 KB reserved per block) and the register budget the build was compiled for
            long per_sm = (long)(SMEM_OPTIN_LIMIT / (spb * per_seq + 1024));
            const long by_regs = mode == 2 ? 2 : 1;  // (narrow / wide: a block of several hundred threads at 70-110 registers)
            if (per_sm > by_regs) per_sm = by_regs;
            if (per_sm < 1) per_sm = 1;
            const long blocks = (nseq + spb - 1) / spb;
            const long waves = (blocks + 148 * per_sm - 1) / (148 * per_sm);
            // (ties stay with the earlier build: at 1024 points wide and deep both need one wave, wide measured 3514 vs 3540 us/slice on config #3)
            if (best < 0 || waves < best_waves) {
                best = mode;
                best_spb = spb;
                best_waves = waves;
            }
        }
        if (best < 0) return fail(c, QPIC_ERR_UNSUPPORTED, "%s: length %d does not fit the shared-memory transform", what, n);
        L->mode = best;
        L->spb = best_spb;
        L->threads = best_spb * e8;
        L->smem = best_spb * (best == 0 ? 2 * one : one);
        L->blocks = (nseq + L->spb - 1) / L->spb;
        return QPIC_OK;
    } else {
        L->spb = 1;
        L->threads = pick_threads(n / 4);
        L->smem = sizeof(double2) * 2 * (size_t)n;
    }
    if (L->threads > lim.max_threads) return fail(c, QPIC_ERR_UNSUPPORTED, "%s: length %d needs more than %d threads per block", what, n, lim.max_threads);
    if (L->smem > lim.max_dyn_smem) return fail(c, QPIC_ERR_UNSUPPORTED, "%s: length %d does not fit the shared-memory transform", what, n);
    L->blocks = (nseq + L->spb - 1) / L->spb;
    return QPIC_OK;
}

// the separable parts of the divisor at src/System_3d.cpp:699, evaluated with the host libm exactly as written there
int make_symbol_tables(qpic_ctx * c) {
    const int ny = c->g.ny, nz = c->g.nz;
    std::vector<double> sy(ny), sz(nz);
    for (int j = 0; j < ny; j++) sy[j] = (2 * cos(2 * M_PI * j / ny) - 2) / c->g.dy / c->g.dy;
    for (int k = 0; k < nz; k++) sz[k] = (2 * cos(2 * M_PI * k / nz) - 2) / c->g.dz / c->g.dz;
    QPIC_TRY(dev_alloc(c, &c->sy, (size_t)ny, false));
    QPIC_TRY(dev_alloc(c, &c->sz, (size_t)nz, false));
    QPIC_CUDA(c, cudaMemcpyAsync(c->sy, sy.data(), sizeof(double) * ny, cudaMemcpyHostToDevice, c->stream));
    QPIC_CUDA(c, cudaMemcpyAsync(c->sz, sz.data(), sizeof(double) * nz, cudaMemcpyHostToDevice, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

// 1 / (-D + sy[j] + sz[min(k, nz-k)]) laid out [k][j] (a column of the spectrum reads it contiguously); the divisor is the
// reference's expression (src/System_3d.cpp:699) evaluated on the host, bin (0,0) is 0 (:688-689)
int get_inv_symbol(qpic_ctx * c, double D, bool full, const double ** out) {
    for (const auto & e : c->inv_symbols) {
        if (e.D == D && e.full == full) {
            *out = e.table;
            return QPIC_OK;
        }
    }
    const int ny = c->g.ny, nz = c->g.nz;
    const int ncols = full ? nz : nz / 2 + 1;
    std::vector<double> h((size_t)ncols * ny);
    for (int k = 0; k < ncols; k++) {
        const int kk = (k <= nz - k) ? k : nz - k;
        for (int j = 0; j < ny; j++) {
            const double multiplier =
                -D + (2 * cos(2 * M_PI * j / ny) - 2) / c->g.dy / c->g.dy + (2 * cos(2 * M_PI * kk / nz) - 2) / c->g.dz / c->g.dz;
            h[(size_t)k * ny + j] = (j == 0 && k == 0) ? 0.0 : 1.0 / multiplier;
        }
    }
    double * d = nullptr;
    QPIC_TRY(dev_alloc(c, &d, h.size(), false));
    QPIC_CUDA(c, cudaMemcpyAsync(d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->inv_symbols.push_back({D, full, d});
    *out = d;
    return QPIC_OK;
}

int check_launch(qpic_ctx * c, const char * what) {
    c->launch_counter++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, QPIC_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    if (c->profiling) {
        cudaEvent_t ev;
        QPIC_CUDA(c, cudaEventCreate(&ev));
        QPIC_CUDA(c, cudaEventRecord(ev, c->stream));
        c->prof_names.push_back(what);
        c->prof_events.push_back(ev);
    }
    return QPIC_OK;
}

// lists of the slice-start positions: values in slot order (identity permutation) ...
CellLists lists_base(qpic_ctx * c) { return CellLists{c->start_b, nullptr, c->P.fyz_b}; }
// lists of the half-step positions: permutation of slots
CellLists lists_mid(qpic_ctx * c) { return CellLists{c->start_m, c->perm_m, c->P.fyz_m}; }

// ---- where things live, single GPU vs one slab of a distributed grid
int row_first(const qpic_ctx * c) { return c->dist.on ? c->dist.v.row0 : 0; }
int row_count(const qpic_ctx * c) { return c->dist.on ? c->dist.v.ny_loc : c->g.ny; }

SpecMap spec_map(const qpic_ctx * c, bool full) {
    SpecMap m = {};
    if (c->dist.on) {
        m.world = c->dist.v.world;
        m.cpr = full ? c->dist.cpr_full : c->dist.cpr_half;
        for (int q = 0; q < m.world; q++) m.base[q] = reinterpret_cast<double2 *>(c->dist.v.base[q] + c->dist.v.lay.cols);
    } else {
        m.world = 1;
        m.cpr = full ? c->g.nz : c->g.nz / 2 + 1;
        m.base[0] = c->spec;
    }
    return m;
}

SpecMap partials_map(const qpic_ctx * c) {
    SpecMap m = {};
    m.world = c->dist.on ? c->dist.v.world : 1;
    for (int q = 0; q < m.world; q++) {
        m.base[q] = c->dist.on ? reinterpret_cast<double2 *>(c->dist.v.base[q] + c->dist.v.lay.partials) : reinterpret_cast<double2 *>(c->partials);
    }
    return m;
}

// ---- exchange steps of a slab-decomposed sweep (no-ops on a single GPU)
int dist_barrier(qpic_ctx * c, cudaStream_t st) {
    if (!c->dist.on) return QPIC_OK;
    QPIC_LAUNCH(k_xchg_barrier, 1, 32, 0, st, c->dist.v, c->ctl);
    return check_launch(c, "k_xchg_barrier");
}

HaloJob halo_job(std::initializer_list<double *> arrays, std::initializer_list<int> slots, std::initializer_list<int> slice3d = {}) {
    HaloJob j = {};
    for (double * a : arrays) j.a[j.n++] = a;
    int i = 0;
    for (int sl : slots) j.slot[i++] = sl;
    i = 0;
    for (int s3 : slice3d) j.slice3d[i++] = s3;
    return j;
}

int halo_blocks(const qpic_ctx * c, const HaloJob & j) {
    int b = blocks_for((size_t)j.n * 2 * DIST_HALO * c->g.nz, 256);
    return b > 592 ? 592 : b;
}

// partial sums that the local particles left in the halo rows -> neighbours; theirs -> added to my edge rows
int dist_reduce(qpic_ctx * c, cudaStream_t st, const HaloJob & j) {
    if (!c->dist.on) return QPIC_OK;
    QPIC_LAUNCH(k_halo_reduce_send, halo_blocks(c, j), 256, 0, st, c->ctl, c->g, c->dist.v, j);
    QPIC_TRY(check_launch(c, "k_halo_reduce_send"));
    QPIC_TRY(dist_barrier(c, st));
    QPIC_LAUNCH(k_halo_reduce_add, halo_blocks(c, j), 256, 0, st, c->ctl, c->g, c->dist.v, j);
    return check_launch(c, "k_halo_reduce_add");
}

// my edge rows of region fields -> the neighbours' halo rows
int dist_fill(qpic_ctx * c, cudaStream_t st, std::initializer_list<int> fields) {
    if (!c->dist.on) return QPIC_OK;
    HaloJob j = {};
    for (int f : fields) {
        j.field_off[j.n] = c->dist.v.lay.field[f];
        j.slot[j.n++] = f;
    }
    QPIC_LAUNCH(k_halo_fill, halo_blocks(c, j), 256, 0, st, c->g, c->dist.v, j);
    QPIC_TRY(check_launch(c, "k_halo_fill"));
    return dist_barrier(c, st);
}

// ---------------------------------------------------------------------------------------------- kernel sequences
//
// Two variants of the slice body exist:
//   FAST     sort-free tile deposits (k_tile_deposit); valid while every half-step displacement stays within +-1 cell.
//            A violation raises Ctl::fast_fail on the device, which turns the rest of the sweep into no-ops; the host
//            notices at its next poll, restores By/Bz of the slice start and redoes the slice on the
//   GENERAL  path: counting sort of the half-step positions + node-centric deposits, valid for any displacement.
// Independent branches of an iteration (corrector gamma + jx deposit  ||  psi_middle solve + jy/jz deposit) are issued
// on two streams joined by events, so they become parallel branches of the captured graph.

cudaEvent_t next_event(qpic_ctx * c) {
    cudaEvent_t e = c->ev_pool[c->ev_next];
    c->ev_next = (c->ev_next + 1) % (int)c->ev_pool.size();
    return e;
}

// work queued on `to` after this call waits for everything queued on `from` so far
int stream_depend(qpic_ctx * c, cudaStream_t from, cudaStream_t to) {
    if (from == to) return QPIC_OK;
    cudaEvent_t e = next_event(c);
    QPIC_CUDA(c, cudaEventRecord(e, from));
    QPIC_CUDA(c, cudaStreamWaitEvent(to, e, 0));
    return QPIC_OK;
}

// count[] (already accumulated by a particle kernel) -> start[], perm[] in canonical order
int enqueue_sort(qpic_ctx * c, cudaStream_t st, const Ctl * ctl, const int * cell, int * start, int * perm, bool with_slot_copies = false) {
    // distributed: bucket n2 collects the entries of emigrated particles; the live count is the start of that bucket
    const int n2 = c->g.n2 + (c->dist.on ? 1 : 0), np = c->P.np;
    int * live_out = (c->dist.on && ctl == c->ctl) ? &c->ctl->np1 : nullptr;
    QPIC_LAUNCH(k_scan_chunk_sums, c->nchunks, SCAN_THREADS, 0, st, ctl, c->count, n2, c->sums);
    QPIC_TRY(check_launch(c, "k_scan_chunk_sums"));
    // few chunks (<= 1024, i.e. grids up to 1024 x 1024): every block of the apply kernel adds up the chunk totals before its own
    // instead of a separate single-block scan kernel
    const int fold = c->nchunks <= 1024 ? 1 : 0;
    if (!fold) {
        QPIC_LAUNCH(k_scan_sums, 1, SCAN_THREADS, 0, st, ctl, c->sums, c->nchunks);
        QPIC_TRY(check_launch(c, "k_scan_sums"));
    }
    QPIC_LAUNCH(k_scan_apply, c->nchunks, SCAN_THREADS, 0, st, ctl, c->count, n2, c->sums, start, c->cursor, live_out, fold);
    QPIC_TRY(check_launch(c, "k_scan_apply"));
    if (np > 0) {
        QPIC_LAUNCH_FLAT(k_sort_fill, blocks_for(np, 256), 256, 0, st, ctl, np, cell, c->cursor, c->tmp);
        QPIC_TRY(check_launch(c, "k_sort_fill"));
        if (with_slot_copies) {
            // (cell == P.cell_p here: the sort of the slice-start positions, whose slot copies ride along)
            QPIC_LAUNCH_FLAT(k_rank_to_slots, blocks_for(np, QPIC_PBLOCK), QPIC_PBLOCK, 0, st, ctl, c->g, c->P, np, (const int *)start, (const int *)c->tmp, perm);
            QPIC_TRY(check_launch(c, "k_particles_to_slots"));
        } else {
            QPIC_LAUNCH_FLAT(k_sort_rank, blocks_for(np, 256), 256, 0, st, ctl, np, cell, start, c->tmp, perm);
            QPIC_TRY(check_launch(c, "k_sort_rank"));
        }
    }
    return QPIC_OK;
}

// columns forward, / symbol, backward over `ncols` columns of spec (leading dimension ld)
int enqueue_columns(qpic_ctx * c, cudaStream_t st, const Ctl * ctl, double D, bool full, bool want_border_sums = false) {
    const int nz = c->g.nz, nzh = nz / 2 + 1;
    const int ncols = full ? nz : nzh;
    const qpic_ctx::FftLaunch & L = full ? c->cols_full : c->cols_half;
    if (c->pow2_y) {
        const double * inv = nullptr;
        QPIC_TRY(get_inv_symbol(c, D, full, &inv));
        // distributed: this rank's block of columns, [ny][cpr] in its exchange region, filled by the peers' row kernels
        const SpecMap m = spec_map(c, full);
        const int col0 = c->dist.on ? (full ? c->dist.col0_full : c->dist.col0_half) : 0;
        const int mine = c->dist.on ? (full ? c->dist.ncols_full : c->dist.ncols_half) : ncols;
        double2 * cols = c->dist.on ? m.base[c->dist.v.rank] : c->spec;
        QPIC_LAUNCH(P2_PICK0(L.mode, k_p2_cols_solve), L.blocks, L.threads, L.smem, st, ctl, c->fy.plan.log2n, c->fy.tw, cols, m.cpr, mine, inv + (size_t)col0 * c->g.ny,
                    want_border_sums ? c->spec_row0 : (double2 *)nullptr, want_border_sums ? c->spec_colsum : (double2 *)nullptr, c->dist.on ? c->cols_ticket : (int *)nullptr,
                    partials_map(c), c->dist.on ? c->dist.v.rank : 0, col0, nz);
    } else {
        QPIC_LAUNCH(k_fft_cols_solve<false>, L.blocks, L.threads, L.smem, st, ctl, c->fy.plan, c->fy.tw, c->spec, ncols, ncols, L.spb, c->sy, c->sz, nz, D);
    }
    return check_launch(c, full ? "k_fft_cols_solve_b" : "k_fft_cols_solve");
}

// real solve: fft_in -> raw (= solution / ny / nz).  With fixed_dst != NULL and both axes on the power-of-two engine the
// potential fix-up (border mean, clamp) is fused into the backward rows and the result goes straight to fixed_dst;
// returns whether that happened through *fused.
int enqueue_real_solve(qpic_ctx * c, cudaStream_t st, const Ctl * ctl, double D, double * fixed_dst = nullptr, bool * fused = nullptr) {
    const int ny = c->g.ny, nz = c->g.nz;
    const qpic_ctx::FftLaunch & R = c->rows_real;
    const bool fuse = fixed_dst != nullptr && c->pow2_y && c->pow2_z;
    if (fused) *fused = fuse;
    if (c->dist.on && !fuse) return fail(c, QPIC_ERR_UNSUPPORTED, "distributed sweeps need the fused power-of-two solve");
    if (c->pow2_z) {
        QPIC_LAUNCH((c->dist.on ? P2_PICK(R.mode, k_p2_rows_r2c, true) : P2_PICK(R.mode, k_p2_rows_r2c, false)), R.blocks, R.threads, R.smem, st, ctl, c->fz.plan.log2n, c->fz.tw, c->fft_in, spec_map(c, false), row_count(c), row_first(c));
    } else {
        QPIC_LAUNCH(k_fft_rows_r2c<false>, R.blocks, R.threads, R.smem, st, ctl, c->fz.plan, c->fz.tw, c->fft_in, c->spec, ny, nz, R.spb);
    }
    QPIC_TRY(check_launch(c, "k_fft_rows_r2c"));
    QPIC_TRY(dist_barrier(c, st));  // the peers' rows of my columns have arrived
    QPIC_TRY(enqueue_columns(c, st, ctl, D, false, fuse));
    QPIC_TRY(dist_barrier(c, st));  // every rank's columns (and border-sum partials) are done
    if (c->pow2_z) {
        QPIC_LAUNCH((c->dist.on ? P2_PICK(R.mode, k_p2_rows_c2r, true) : P2_PICK(R.mode, k_p2_rows_c2r, false)), R.blocks, R.threads, R.smem, st, ctl, c->fz.plan.log2n, c->fz.tw, spec_map(c, false), fuse ? fixed_dst : c->raw, row_count(c),
                    row_first(c), ny, c->inv_scale, fuse ? 1 : 0, c->dist.on ? reinterpret_cast<const double *>(c->dist.region + c->dist.v.lay.partials) : c->partials,
                    c->prm.psi_threshold - 1, (fuse && !c->dist.on) ? (const double2 *)c->spec_row0 : (const double2 *)nullptr,
                    (fuse && !c->dist.on) ? (const double2 *)c->spec_colsum : (const double2 *)nullptr);
    } else {
        QPIC_LAUNCH(k_fft_rows_c2r<false>, R.blocks, R.threads, R.smem, st, ctl, c->fz.plan, c->fz.tw, c->spec, c->raw, ny, nz, R.spb, c->inv_scale);
    }
    QPIC_TRY(check_launch(c, "k_fft_rows_c2r"));
    return QPIC_OK;
}

// complex solve of (re, im) = (by, bz) style pairs: spec must already hold the forward row transforms
int enqueue_complex_columns(qpic_ctx * c, cudaStream_t st, const Ctl * ctl, double D) { return enqueue_columns(c, st, ctl, D, true); }

// backward rows of a complex solve into (by, bz) with the stop test
int enqueue_b_backward(qpic_ctx * c, cudaStream_t st, Ctl * ctl, double * by, double * bz, double threshold, int iteration) {
    const qpic_ctx::FftLaunch & R = c->rows_cplx;
    if (c->pow2_z) {
        QPIC_LAUNCH((c->dist.on ? P2_PICK(R.mode, k_p2_rows_b_backward, true) : P2_PICK(R.mode, k_p2_rows_b_backward, false)), R.blocks, R.threads, R.smem, st, ctl, c->fz.plan.log2n, c->fz.tw, c->g, spec_map(c, true), by, bz, threshold, iteration,
                    c->inv_scale, row_first(c), row_count(c));
    } else {
        QPIC_LAUNCH(k_fft_rows_b_backward<false>, R.blocks, R.threads, R.smem, st, ctl, c->fz.plan, c->fz.tw, c->g, c->spec, by, bz, threshold, iteration,
                    R.spb, c->inv_scale);
    }
    return check_launch(c, "k_fft_rows_b_backward");
}

// ---- tile deposit jobs (FAST path)

// distributed: the tile rows the rank's particles can reach = own rows + one tile row on either side
dim3 tile_grid(qpic_ctx * c) { return dim3((c->g.nz + TILE - 1) / TILE, c->dist.on ? c->dist.v.ny_loc / TILE + 2 : (c->g.ny + TILE - 1) / TILE); }

TileOutput tile_output(const double * val, double sign, int bg_mode, const double * bg3, double * out, int out_3d) {
    TileOutput o;
    o.val = val;
    o.sign = sign;
    o.bg_mode = bg_mode;
    o.bg3 = bg3;
    o.out = out;
    o.out_3d = out_3d;
    return o;
}

TileJob tile_job(qpic_ctx * c) {
    TileJob j;
    std::memset(&j, 0, sizeof(j));
    j.jxb3 = c->f3[QPIC_F3_JX_BUNCH];
    j.rhob3 = c->f3[QPIC_F3_RHO_BUNCH];
    j.rho_ion = c->f2[QPIC_F2_RHO_ION];
    j.tile_rows = (c->g.ny + TILE - 1) / TILE;
    j.tile_row0 = c->dist.on ? (c->dist.v.row0 / TILE - 1 + j.tile_rows) % j.tile_rows : 0;
    j.own_row0 = row_first(c);
    j.own_rows = row_count(c);
    if (c->jyjz_lists) {
        j.bin_cnt = c->bin_cnt;
        j.bin_lst = c->bin_lst;
        // rows of the target cells of the launched tile rows: j0 - 1 ... j0 + TILE - 1 each
        j.list_row0 = c->dist.on ? (j.tile_row0 * TILE - 1 + c->g.ny) % c->g.ny : 0;
        j.list_rows = c->dist.on ? (c->dist.v.ny_loc / TILE + 2) * TILE + 1 : c->g.ny;
    }
    // node rows of the launched tile rows
    j.cell_row0 = c->dist.on ? j.tile_row0 * TILE : 0;
    j.cell_rows = c->dist.on ? (c->dist.v.ny_loc / TILE + 2) * TILE : c->g.ny;
    return j;
}

// deposits at (y,z)
int launch_tile(qpic_ctx * c, cudaStream_t st, TileJob & j, bool final_pass, const char * name) {
    const dim3 grid = tile_grid(c), block(TILE_ACC, TILE_ACC);
    if (j.nout == 1) {
        auto kern = final_pass ? k_tile_deposit<1, true> : k_tile_deposit<1, false>;
        QPIC_LAUNCH(kern, grid, block, 0, st, c->ctl, c->g, c->start_b, c->P, j);
    } else if (j.nout == 2) {
        // rho + susceptibility after the B iterations: the values are formed in the deposit (the predictor kernel does not store them)
        auto kern = final_pass ? k_tile_deposit<2, true, true> : k_tile_deposit<2, false>;
        QPIC_LAUNCH(kern, grid, block, 0, st, c->ctl, c->g, c->start_b, c->P, j);
    } else {
        auto kern = final_pass ? k_tile_deposit<3, true> : k_tile_deposit<3, false>;
        QPIC_LAUNCH(kern, grid, block, 0, st, c->ctl, c->g, c->start_b, c->P, j);
    }
    return check_launch(c, name);
}

// psi source from positions (y,z) [at_mid = 0] or the half-step positions [at_mid = 1] -> fft_in
// with the published lists the psi_middle source deposit and the jy/jz deposit treat the particles that stay in their cell directly
// and exchange only the movers (k_tile_stay_deposit / k_tile_assemble, k_tile_cell_deposit)
bool stayers_direct(const qpic_ctx * c) { return c->jyjz_lists; }
// slabs: target rows whose deposits (corner rows j, j + 1; the staggered jy rows j - 1 .. j + 1) stay inside own rows + DIST_HALO halo rows
void far_rows(const qpic_ctx * c, int * row_lo, int * row_n) {
    *row_lo = 0;
    *row_n = 0;
    if (c->dist.on) {
        *row_lo = (c->dist.v.row0 - DIST_HALO + 1 + c->g.ny) % c->g.ny;
        *row_n = c->dist.v.ny_loc + 2 * DIST_HALO - 2;
    }
}
// far-mover lists of B iteration c->cur_it (only with the stayers-direct tile kernels on a single GPU and without the predictor-side deposit,
// whose own lists take every mover)
FarMovers far_view(const qpic_ctx * c, bool fast) {
    FarMovers f = {nullptr, nullptr, nullptr, 0, 0};
    if (fast && c->far_movers && c->far_cnt[0] && !(c->predict_deposit && c->stay.corn)) {
        f.cnt = c->far_cnt[c->cur_it & 1];
        f.lst = c->far_lst;
        f.clear = c->far_cnt[(c->cur_it + 1) & 1];
        far_rows(c, &f.row_lo, &f.row_n);
    }
    return f;
}
bool predict_deposit_on(const qpic_ctx * c) { return c->predict_deposit && stayers_direct(c) && c->stay.corn != nullptr; }
// ... and the jx deposit of the iteration with it (k_particles_predict_dep<., true> + k_tile_assemble<true>)
bool fold_jx_on(const qpic_ctx * c) { return predict_deposit_on(c) && c->fold_jx && c->stay.xcorn != nullptr; }
// ... and, in the last pass of a slice, rho and susceptibility(i) (k_particles_predict_dep<., 2> + k_tile_fold_rho_chi)
bool fold_rc_on(const qpic_ctx * c) { return fold_jx_on(c) && c->stay.rcorn != nullptr; }

int enqueue_tile_charge(qpic_ctx * c, cudaStream_t st, int at_mid) {
    TileJob j = tile_job(c);
    j.nout = 1;
    j.o[0] = tile_output(c->P.nw_s, -1.0, 1, nullptr, c->fft_in, 0);
    if (!at_mid) {
        if (predict_deposit_on(c)) j.clear_cnt = c->stay_cnt[0];  // the first predictor pass of the slice counts its movers there
        else if (c->far_movers && c->far_cnt[0]) j.clear_cnt = c->far_cnt[0];  // ... or its far movers
        return launch_tile(c, st, j, false, "k_tile_deposit_charge");
    }
    const dim3 block(TILE_ACC + 2, TILE_ACC + 2);
    if (predict_deposit_on(c)) {
        // the predictor of this iteration left the stayers' corner sums and the mover lists (k_particles_predict_dep)
        StayOut so = c->stay;
        so.cnt = c->stay_cnt[c->cur_it & 1];
        if (fold_jx_on(c)) {
            j.nout = 2;
            j.o[1] = tile_output(c->P.vjx, 1.0, 2, c->f3[QPIC_F3_JX_BUNCH], c->f2[QPIC_F2_JX], 0);
            QPIC_LAUNCH(k_tile_assemble<true>, tile_grid(c), dim3(TILE_ACC, TILE_ACC), 0, st, c->ctl, c->g, c->start_b, c->P, j, so, c->stay_cnt[(c->cur_it + 1) & 1]);
        } else {
            QPIC_LAUNCH(k_tile_assemble<false>, tile_grid(c), dim3(TILE_ACC, TILE_ACC), 0, st, c->ctl, c->g, c->start_b, c->P, j, so, c->stay_cnt[(c->cur_it + 1) & 1]);
        }
        return check_launch(c, "k_tile_deposit_charge_mid");
    }
    if (!stayers_direct(c)) return fail(c, QPIC_ERR_STATE, "the sort-free deposit path needs the mover lists (force_general_deposit is set)");
    // stayers from registers, movers through the lists; publishes the movers for k_tile_cell_deposit
    QPIC_LAUNCH(k_tile_stay_deposit, tile_grid(c), block, tile_stay_smem_bytes(), st, c->ctl, c->g, c->start_b, c->P, j, far_view(c, true));
    return check_launch(c, "k_tile_deposit_charge_mid");
}

// deposits at (y,z): jx and / or rho + susceptibility(i)
int enqueue_tile_base_currents(qpic_ctx * c, cudaStream_t st, bool with_jx, bool with_rho_chi, bool final_pass) {
    TileJob j = tile_job(c);
    int n = 0;
    if (with_jx) j.o[n++] = tile_output(c->P.vjx, 1.0, 2, c->f3[QPIC_F3_JX_BUNCH], c->f2[QPIC_F2_JX], 0);
    if (with_rho_chi) {
        j.o[n++] = tile_output(c->P.vrho, 1.0, 2, c->f3[QPIC_F3_RHO_BUNCH], c->f2[QPIC_F2_RHO], 0);
        j.o[n++] = tile_output(c->P.vchi, 1.0, 0, nullptr, c->f3[QPIC_F3_SUSCEPTIBILITY], 1);
    }
    j.nout = n;
    if (with_rho_chi && !with_jx && final_pass) j.rc_folded = c->rc_folded_now;
    return launch_tile(c, st, j, final_pass, with_jx ? (with_rho_chi ? "k_tile_deposit_jx_rho_chi" : "k_tile_deposit_jx") : "k_tile_deposit_rho_chi");
}

// jy_next (staggered in y) and jz_next (staggered in z) at the half-step positions
int enqueue_tile_jy_jz(qpic_ctx * c, cudaStream_t st) {
    TileJob j = tile_job(c);
    j.nout = 2;
    // values n py_next / (1 + psi_middle@particle) are formed inside the deposit (no k_particles_currents_yz on this path)
    // (the kernels read py_next / pz_next from the paired slot array P.pyzn; `val` is unused by this job)
    j.o[0] = tile_output(nullptr, 1.0, 0, nullptr, c->f2[QPIC_F2_JY_NEXT], 0);
    j.o[1] = tile_output(nullptr, 1.0, 0, nullptr, c->f2[QPIC_F2_JZ_NEXT], 0);
    j.psi_mid = c->f2[QPIC_F2_PSI_MIDDLE];
    j.nw = c->P.nw_s;
    if (!stayers_direct(c)) return fail(c, QPIC_ERR_STATE, "the sort-free deposit path needs the mover lists (force_general_deposit is set)");
    // the mover lists of the half-step positions were published by the psi_middle source kernel of this iteration
    j.b_src = c->b_src;
    j.b_by = c->f2[QPIC_F2_BY_CUR];
    j.b_bz = c->f2[QPIC_F2_BZ_CUR];
    j.b_jx = c->f2[QPIC_F2_JX];
    j.b_jy = c->f2[QPIC_F2_JY];
    j.b_jz = c->f2[QPIC_F2_JZ];
    j.b_D = c->prm.magnetic_field_D;
    // one visit per particle
    if (c->dense_cells) {
        constexpr int TZ = 32;
        const dim3 grid((c->g.nz + TZ - 1) / TZ, (j.cell_rows + LIST_TY - 1) / LIST_TY);
        QPIC_LAUNCH(k_tile_cell_deposit<TZ>, grid, dim3(TZ + 2, LIST_BY), sizeof(double) * 12 * LIST_BY * (TZ + 2), st, c->ctl, c->g, c->start_b, c->P, j);
    } else {
        constexpr int TZ = 16;
        const dim3 grid((c->g.nz + TZ - 1) / TZ, (j.cell_rows + LIST_TY - 1) / LIST_TY);
        QPIC_LAUNCH(k_tile_cell_deposit<TZ>, grid, dim3(TZ + 2, LIST_BY), sizeof(double) * 12 * LIST_BY * (TZ + 2), st, c->ctl, c->g, c->start_b, c->P, j);
    }
    return check_launch(c, "k_tile_deposit_jy_jz");
}

// ---- potentials

int enqueue_fixup(qpic_ctx * c, cudaStream_t st, double * dst) {
    int fix_blocks = blocks_for(c->g.n2, FIX_THREADS * 4);
    if (fix_blocks < 1) fix_blocks = 1;
    QPIC_LAUNCH(k_potential_fixup, fix_blocks, FIX_THREADS, 0, st, c->ctl, c->g, c->raw, dst, c->prm.psi_threshold - 1);
    return check_launch(c, "k_potential_fixup");
}

// psi (at_mid = 0) or psi_middle (at_mid = 1): source deposit, solve, normalise + clamp (src/System_3d.cpp:342-365, :446-469)
int enqueue_potential(qpic_ctx * c, cudaStream_t st, bool fast, int at_mid, double * dst) {
    if (fast) {
        QPIC_TRY(enqueue_tile_charge(c, st, at_mid));
    } else {
        CellLists L = at_mid ? lists_mid(c) : lists_base(c);
        QPIC_LAUNCH_FLAT(k_deposit_charge, blocks_for(c->g.n2, 128), 128, 0, st, c->ctl, c->g, L, c->P.nw_s, c->f2[QPIC_F2_RHO_ION], c->f3[QPIC_F3_JX_BUNCH],
                         c->f3[QPIC_F3_RHO_BUNCH], 1, c->fft_in);
        QPIC_TRY(check_launch(c, "k_deposit_charge"));
    }
    // distributed: at_mid sources are reduced by the caller together with jx (one exchange)
    if (!at_mid) QPIC_TRY(dist_reduce(c, st, halo_job({c->fft_in}, {RB_CHARGE})));
    bool fused = false;
    QPIC_TRY(enqueue_real_solve(c, st, c->ctl, 0.0, dst, &fused));
    if (fused) {
        if (!at_mid) QPIC_TRY(dist_fill(c, st, {DF_PSI}));
        return QPIC_OK;
    }
    return enqueue_fixup(c, st, dst);
}

// deposits at (y,z) on either path
int enqueue_base_currents(qpic_ctx * c, cudaStream_t st, bool fast, bool with_jx, bool with_rho_chi, bool final_pass) {
    if (fast) return enqueue_tile_base_currents(c, st, with_jx, with_rho_chi, final_pass);
    const int blocks = blocks_for(c->g.n2, 128);
    auto kern = with_jx ? (with_rho_chi ? k_deposit_jx_rho_chi<true, true> : k_deposit_jx_rho_chi<true, false>) : k_deposit_jx_rho_chi<false, true>;
    QPIC_LAUNCH_FLAT(kern, blocks, 128, 0, st, c->ctl, c->g, lists_base(c), c->P.vjx, c->P.vrho, c->P.vchi, c->f3[QPIC_F3_JX_BUNCH], c->f3[QPIC_F3_RHO_BUNCH],
                     c->f2[QPIC_F2_JX], c->f2[QPIC_F2_RHO], c->f3[QPIC_F3_SUSCEPTIBILITY], final_pass ? 1 : 0);
    return check_launch(c, with_jx ? (with_rho_chi ? "k_deposit_jx_rho_chi" : "k_deposit_jx") : "k_deposit_rho_chi");
}

// K1..K6 of SURVEY.md section 3.2: psi, gamma/px (src/System_3d.cpp:334-378).  with_deposit: also rho/jx/susceptibility
// (:381-397) -- needed only when no B iteration follows (M = 0, or slice 0), every iteration redeposits them.
int enqueue_slice_head(qpic_ctx * c, bool fast, bool want_invariants, bool with_deposit) {
    const int np = c->P.np;
    cudaStream_t st = c->stream;
    QPIC_TRY(enqueue_potential(c, st, fast, 0, c->f2[QPIC_F2_PSI]));
    // with B iterations to follow, the first predictor pass gathers the invariants and forms the initial gamma itself (predict_inputs<true>)
    if (np > 0 && !(want_invariants && c->fuse_gamma)) {
        QPIC_LAUNCH_FLAT(k_particles_gamma, blocks_for(np, QPIC_PBLOCK), QPIC_PBLOCK, 0, st, c->ctl, c->g, c->P, c->perm_b, c->f3[QPIC_F3_ASQR], c->f2[QPIC_F2_PSI],
                         want_invariants ? 1 : 0);
        QPIC_TRY(check_launch(c, "k_particles_gamma"));
    }
    if (with_deposit) {
        QPIC_TRY(enqueue_base_currents(c, st, fast, true, true, false));
        QPIC_TRY(dist_reduce(c, st, halo_job({c->f2[QPIC_F2_JX], c->f2[QPIC_F2_RHO], c->f3[QPIC_F3_SUSCEPTIBILITY]}, {RB_JX, RB_RHO, RB_CHI}, {0, 0, 1})));
    }
    return QPIC_OK;
}

SliceOut slice_out(qpic_ctx * c) {
    SliceOut o;
    o.ex3 = c->f3[QPIC_F3_EX];
    o.ey3 = c->f3[QPIC_F3_EY];
    o.ez3 = c->f3[QPIC_F3_EZ];
    o.by3 = c->f3[QPIC_F3_BY];
    o.bz3 = c->f3[QPIC_F3_BZ];
    o.psi3 = c->f3[QPIC_F3_PSI];
    o.rho3 = c->f3[QPIC_F3_RHO];
    o.jx3 = c->f3[QPIC_F3_JX];
    o.jy3 = c->f3[QPIC_F3_JY];
    o.jz3 = c->f3[QPIC_F3_JZ];
    return o;
}

// z0 of the Fields_xy cuts and the two planes / weight it falls between on a grid whose z-origin is oz
double xy_z0(const qpic_ctx * c) { return c->prm.xy_z0 < 0 ? 0.5 * c->prm.lz : c->prm.xy_z0; }

XyCut xy_cut(const qpic_ctx * c, double oz) {
    XyCut x;
    const double t = (xy_z0(c) - oz) / c->g.dz;  // src/array_utils.cpp:14-15
    int k0 = (int)std::floor(t);
    x.frac = t - k0;
    k0 = ((k0 % c->g.nz) + c->g.nz) % c->g.nz;
    x.k0 = k0;
    x.k1 = k0 + 1 >= c->g.nz ? 0 : k0 + 1;
    return x;
}

int enqueue_end_of_slice(qpic_ctx * c, int b_is_zero, int roll, int ignore_stop) {
    if (c->xy[0]) {
        XyLines o = {c->xy[0], c->xy[1], c->xy[2], c->xy[3], c->xy[4]};
        QPIC_LAUNCH_FLAT(k_xy_lines, blocks_for(row_count(c), 128), 128, 0, c->stream, c->ctl, c->g, o, xy_cut(c, 0.0), xy_cut(c, 0.5 * c->g.dz), c->f2[QPIC_F2_PSI],
                         c->f2[QPIC_F2_RHO], c->f2[QPIC_F2_JX], c->f2[QPIC_F2_JY], c->f2[QPIC_F2_JZ], ignore_stop, row_first(c), row_count(c));
        QPIC_TRY(check_launch(c, "k_xy_lines"));
    }
    QPIC_LAUNCH(k_end_of_slice, blocks_for((size_t)row_count(c) * c->g.nz, 256), 256, 0, c->stream, c->ctl, c->g, slice_out(c), c->f2[QPIC_F2_PSI],
                     c->f2[QPIC_F2_PSI_PREV], c->f2[QPIC_F2_BY_CUR], c->f2[QPIC_F2_BZ_CUR], c->f2[QPIC_F2_RHO], c->f2[QPIC_F2_JX], c->f2[QPIC_F2_JY],
                     c->f2[QPIC_F2_JZ], c->f2[QPIC_F2_JY_NEXT], c->f2[QPIC_F2_JZ_NEXT], b_is_zero, roll, ignore_stop, row_first(c), row_count(c), c->ctl,
                     roll ? c->eos_ticket : (int *)nullptr);
    return check_launch(c, "k_end_of_slice");
}

MigrateView migrate_view(qpic_ctx * c) {
    MigrateView m = {};
    if (!c->dist.on) return m;
    const DistView & v = c->dist.v;
    m.enabled = 1;
    m.row0 = v.row0;
    m.ny_loc = v.ny_loc;
    m.cap = v.mig_cap;
    // box 0 of a rank collects what its upper neighbour sends down, box 1 what its lower neighbour sends up
    m.out[0] = reinterpret_cast<double *>(v.base[v.below] + v.lay.inbox);
    m.out[1] = reinterpret_cast<double *>(v.base[v.above] + v.lay.inbox) + (size_t)6 * v.mig_cap;
    m.out_count = c->dist.out_count;
    m.error = &c->ctl->dist_error;
    return m;
}

// one full xi-loop iteration for a slice i >= 1 (src/System_3d.cpp:328-619)
int enqueue_slice_full(qpic_ctx * c, bool fast) {
    const int np = c->P.np, n2 = c->g.n2;
    const int M = c->prm.magnetic_field_iterations;
    cudaStream_t sa = c->stream;
    // per-kernel timing wants one serial stream; so do distributed sweeps (all ranks must issue their exchange barriers in
    // the same order)
    cudaStream_t sb = (c->profiling || c->dist.on) ? c->stream : c->stream2;
    if (c->dist.on && !fast) return fail(c, QPIC_ERR_UNSUPPORTED, "distributed sweeps run on the sort-free deposit path only");
    QPIC_TRY(enqueue_slice_head(c, fast, M > 0, M == 0));
    for (int it = 0; it < M; it++) {
        c->cur_it = it;
        if (np > 0 && fast && predict_deposit_on(c)) {
            StayOut so = c->stay;
            so.cnt = c->stay_cnt[it & 1];
            const bool first = it == 0 && c->fuse_gamma;
            auto predict = first ? k_particles_predict_dep<true> : k_particles_predict_dep<false>;
            if (fold_jx_on(c)) {
                const bool last_rc = fold_rc_on(c) && it == M - 1;
                // (dense cells: the per-cell sums leave the warp through shared memory, see k_particles_predict_dep)
                if (c->dense_cells) {
                    predict = last_rc ? (first ? k_particles_predict_dep<true, 2, true> : k_particles_predict_dep<false, 2, true>)
                                      : (first ? k_particles_predict_dep<true, 1, true> : k_particles_predict_dep<false, 1, true>);
                } else {
                    predict = last_rc ? (first ? k_particles_predict_dep<true, 2> : k_particles_predict_dep<false, 2>)
                                      : (first ? k_particles_predict_dep<true, 1> : k_particles_predict_dep<false, 1>);
                }
            }
            QPIC_LAUNCH(predict, blocks_for(np, QPIC_PBLOCK), QPIC_PBLOCK, 0, sa, c->ctl, c->g, c->P, c->f2[QPIC_F2_BY_CUR], c->f2[QPIC_F2_BZ_CUR], (it == M - 1) ? 1 : 0, so,
                        (const double *)c->f3[QPIC_F3_ASQR], (const double *)c->f2[QPIC_F2_PSI]);
            QPIC_TRY(check_launch(c, "k_particles_predict"));
        } else if (np > 0) {
            auto predict = fast ? k_particles_predict<true> : k_particles_predict<false>;
            if (it == 0 && c->fuse_gamma) predict = fast ? k_particles_predict<true, true> : k_particles_predict<false, true>;
            QPIC_LAUNCH_FLAT(predict, blocks_for(np, QPIC_PBLOCK), QPIC_PBLOCK, 0, sa, c->ctl, c->g, c->P, c->f2[QPIC_F2_BY_CUR], c->f2[QPIC_F2_BZ_CUR], c->count,
                             (!fast || it == M - 1) ? 1 : 0, far_view(c, fast && stayers_direct(c)), (const double *)c->f3[QPIC_F3_ASQR], (const double *)c->f2[QPIC_F2_PSI]);
            QPIC_TRY(check_launch(c, "k_particles_predict"));
        }
        // branch B: the jx deposit at (y,z) needs only the corrector values the predictor kernel has just stored.  (Forking it after
        // the psi_middle source deposit instead, so that it overlaps the three transform kernels of the solve rather than the other
        // tile kernel, was measured: 0.696 -> 0.740 ms/slice -- a tile deposit next to the few-warp transform kernels delays them.)
        // (With the predictor-side deposit the jx sums were formed by the predictor and k_tile_assemble writes jx: no branch B.)
        const bool jx_folded = np > 0 && fast && fold_jx_on(c);
        if (!jx_folded) {
            QPIC_TRY(stream_depend(c, sa, sb));
            QPIC_TRY(enqueue_base_currents(c, sb, fast, true, false, false));
        }
        // branch A: psi_middle, then jy_next / jz_next
        if (!fast) QPIC_TRY(enqueue_sort(c, sa, c->ctl, c->P.cell_m, c->start_m, c->perm_m));
        if (c->dist.on) {
            // source deposit, then ONE exchange for the psi_middle source and jx, then the solve
            QPIC_TRY(enqueue_tile_charge(c, sa, 1));
            QPIC_TRY(dist_reduce(c, sa, halo_job({c->fft_in, c->f2[QPIC_F2_JX]}, {RB_CHARGE, RB_JX})));
            bool fused = false;
            QPIC_TRY(enqueue_real_solve(c, sa, c->ctl, 0.0, c->f2[QPIC_F2_PSI_MIDDLE], &fused));
            QPIC_TRY(dist_fill(c, sa, {DF_PSI_MIDDLE, DF_JX}));
        } else {
            QPIC_TRY(enqueue_potential(c, sa, fast, 1, c->f2[QPIC_F2_PSI_MIDDLE]));
        }
        if (fast) {
            if (c->jyjz_lists && c->b_src && !jx_folded) QPIC_TRY(stream_depend(c, sb, sa));  // the deposit's epilogue reads jx (branch B)
            QPIC_TRY(enqueue_tile_jy_jz(c, sa));
            QPIC_TRY(dist_reduce(c, sa, halo_job({c->f2[QPIC_F2_JY_NEXT], c->f2[QPIC_F2_JZ_NEXT]}, {RB_JY, RB_JZ})));
        } else {
            if (np > 0) {
                QPIC_LAUNCH_FLAT(k_particles_currents_yz, blocks_for(np, QPIC_PBLOCK), QPIC_PBLOCK, 0, sa, c->ctl, c->g, c->P, c->f2[QPIC_F2_PSI_MIDDLE]);
                QPIC_TRY(check_launch(c, "k_particles_currents_yz"));
            }
            QPIC_LAUNCH_FLAT(k_deposit_jy_jz, blocks_for(n2, 128), 128, 0, sa, c->ctl, c->g, lists_mid(c), c->P.vjy, c->P.vjz, c->f2[QPIC_F2_JY_NEXT],
                             c->f2[QPIC_F2_JZ_NEXT]);
            QPIC_TRY(check_launch(c, "k_deposit_jy_jz"));
        }
        if (!jx_folded) QPIC_TRY(stream_depend(c, sb, sa));
        // By + i Bz in one complex solve
        {
            const qpic_ctx::FftLaunch & R = c->rows_cplx;
            if (c->pow2_z) {
                const bool pre = fast && c->jyjz_lists && c->b_src;
                auto kern = c->dist.on ? P2_PICK(R.mode, k_p2_rows_b_forward, true, false)
                                       : (pre ? P2_PICK(R.mode, k_p2_rows_b_forward, false, true) : P2_PICK(R.mode, k_p2_rows_b_forward, false, false));
                QPIC_LAUNCH(kern, R.blocks, R.threads, R.smem, sa, c->ctl, c->fz.plan.log2n, c->fz.tw, c->g, c->prm.magnetic_field_D, c->f2[QPIC_F2_BY_CUR],
                            c->f2[QPIC_F2_BZ_CUR], c->f2[QPIC_F2_JX], c->f2[QPIC_F2_JY], c->f2[QPIC_F2_JZ], c->f2[QPIC_F2_JY_NEXT], c->f2[QPIC_F2_JZ_NEXT],
                            (const double2 *)c->b_src, spec_map(c, true), row_first(c), row_count(c));
            } else {
                QPIC_LAUNCH(k_fft_rows_b_forward<false>, R.blocks, R.threads, R.smem, sa, c->ctl, c->fz.plan, c->fz.tw, c->g, c->prm.magnetic_field_D,
                            c->f2[QPIC_F2_BY_CUR], c->f2[QPIC_F2_BZ_CUR], c->f2[QPIC_F2_JX], c->f2[QPIC_F2_JY], c->f2[QPIC_F2_JZ], c->f2[QPIC_F2_JY_NEXT],
                            c->f2[QPIC_F2_JZ_NEXT], c->spec, R.spb);
            }
            QPIC_TRY(check_launch(c, "k_fft_rows_b_forward"));
            QPIC_TRY(dist_barrier(c, sa));
            QPIC_TRY(enqueue_complex_columns(c, sa, c->ctl, c->prm.magnetic_field_D));
            QPIC_TRY(dist_barrier(c, sa));
            QPIC_TRY(enqueue_b_backward(c, sa, c->ctl, c->f2[QPIC_F2_BY_CUR], c->f2[QPIC_F2_BZ_CUR], c->prm.threshold_B, it));
            QPIC_TRY(dist_fill(c, sa, {DF_BY, DF_BZ}));
        }
    }
    // branch B: rho and susceptibility(i) from the last executed iteration (see k_deposit_jx_rho_chi), then the fields of this slice
    // + the psi_prev / jy / jz roll;  branch A: push + cell lists of the new positions.  The deposit reads the cell lists and the
    // slot arrays of (y, z): the push only writes particle-id arrays, so the two overlap (the deposit is latency bound, the push
    // bandwidth bound); the cell sort below, which rewrites the lists, waits for the deposit.
    QPIC_TRY(stream_depend(c, sa, sb));
    cudaEvent_t deposit_done = nullptr;
    if (M > 0) {
        const bool rc_folded = np > 0 && fast && fold_rc_on(c);
        if (rc_folded) {
            // the last predictor pass left the per-cell sums; the tile kernel below then only serves a slice that the |B| stop condition
            // ended before that pass
            TileJob j = tile_job(c);
            j.nout = 2;
            j.o[0] = tile_output(c->P.vrho, 1.0, 2, c->f3[QPIC_F3_RHO_BUNCH], c->f2[QPIC_F2_RHO], 0);
            j.o[1] = tile_output(c->P.vchi, 1.0, 0, nullptr, c->f3[QPIC_F3_SUSCEPTIBILITY], 1);
            QPIC_LAUNCH(k_tile_fold_rho_chi, tile_grid(c), dim3(TILE_ACC, TILE_ACC), 0, sb, c->ctl, c->g, c->start_b, j, c->stay, M - 1);
            QPIC_TRY(check_launch(c, "k_tile_fold_rho_chi"));
        }
        c->rc_folded_now = rc_folded ? M : 0;
        QPIC_TRY(enqueue_base_currents(c, sb, fast, false, true, true));
        c->rc_folded_now = 0;
        QPIC_TRY(dist_reduce(c, sb, halo_job({c->f2[QPIC_F2_RHO], c->f3[QPIC_F3_SUSCEPTIBILITY]}, {RB_RHO, RB_CHI}, {0, 1})));
        if (sb != sa) {
            deposit_done = next_event(c);
            QPIC_CUDA(c, cudaEventRecord(deposit_done, sb));
        }
    }
    {
        cudaStream_t keep = c->stream;
        c->stream = sb;
        int rc = enqueue_end_of_slice(c, 0, 1, 0);
        c->stream = keep;
        QPIC_TRY(rc);
    }
    if (np > 0) {
        QPIC_LAUNCH_FLAT(k_particles_advance, blocks_for(np, QPIC_PBLOCK), QPIC_PBLOCK, 0, sa, c->ctl, c->g, c->P, c->dist.on ? (const int *)nullptr : c->perm_b,
                         c->f2[QPIC_F2_PSI_MIDDLE], c->count, migrate_view(c), c->track_gamma_px ? 1 : 0);
        QPIC_TRY(check_launch(c, "k_particles_advance"));
    }
    if (c->dist.on) {
        // emigrants are in the neighbours' inboxes: publish the counts, meet, append the arrivals to the id arrays
        QPIC_LAUNCH(k_migrate_publish, 1, 32, 0, sa, c->dist.v, c->dist.out_count);
        QPIC_TRY(check_launch(c, "k_migrate_publish"));
        QPIC_TRY(dist_barrier(c, sa));
        QPIC_LAUNCH(k_migrate_unpack, 148, 256, 0, sa, c->ctl, c->g, c->dist.v, c->P, c->P.np, c->count);
        QPIC_TRY(check_launch(c, "k_migrate_unpack"));
    }
    if (deposit_done) QPIC_CUDA(c, cudaStreamWaitEvent(sa, deposit_done, 0));
    QPIC_TRY(enqueue_sort(c, sa, c->ctl, c->P.cell_p, c->start_b, c->perm_b, true));
    // (the last block of k_end_of_slice has moved Ctl::slice on)
    QPIC_TRY(stream_depend(c, sb, sa));
    return QPIC_OK;
}

// (re)build the cell lists of (y, z) from scratch
int enqueue_rebuild_base_lists(qpic_ctx * c) {
    // distributed: P.np is the capacity of the arrays, the live entries are counted on the device (Ctl::np1)
    const int np = c->P.np, live = c->dist.on ? c->dist.live0 : np;
    Ctl * ctl = c->dist.on ? c->ctl : c->ctl_scratch;
    QPIC_CUDA(c, cudaMemsetAsync(c->count, 0, sizeof(int) * (c->g.n2 + 1), c->stream));
    if (live > 0) {
        QPIC_LAUNCH_FLAT(k_locate_count, blocks_for(live, 256), 256, 0, c->stream, live, c->g, c->P.y, c->P.z, c->P.cell_p, c->P.fy_p, c->P.fz_p, c->count);
        QPIC_TRY(check_launch(c, "k_locate_count"));
    }
    QPIC_TRY(enqueue_sort(c, c->stream, ctl, c->P.cell_p, c->start_b, c->perm_b, true));
    c->lists_dirty = false;
    return QPIC_OK;
}

int alloc_particles(qpic_ctx * c, size_t count) {
    Particles & P = c->P;
    P.np = (int)count;
    double ** dbl[] = {&P.y, &P.z, &P.py, &P.pz, &P.nw, &P.gamma, &P.px, &P.fy_p, &P.fz_p, &P.y_s, &P.z_s, &P.nw_s, &P.gam, &P.pxs, &P.vjx, &P.vrho, &P.vchi, &P.vjy, &P.vjz};
    for (double ** p : dbl) QPIC_TRY(dev_alloc(c, p, count));
    double2 ** pairs[] = {&P.fyz_b, &P.pyz_s, &P.pyzn, &P.fyz_m, &P.qa, &P.t1, &P.dp};
    for (double2 ** p : pairs) QPIC_TRY(dev_alloc(c, p, count));
    QPIC_TRY(dev_alloc(c, &P.cell_p, count));
    QPIC_TRY(dev_alloc(c, &P.cell_b, count));
    QPIC_TRY(dev_alloc(c, &P.cell_m, count));
    QPIC_TRY(dev_alloc(c, &P.off_m, count));
    if (c->dist.on) QPIC_TRY(dev_alloc(c, &P.akey, 2 * (size_t)c->dist.v.mig_cap));  // sort keys of a slice's arrivals
    QPIC_TRY(dev_alloc(c, &c->perm_b, count));
    QPIC_TRY(dev_alloc(c, &c->perm_m, count));
    QPIC_TRY(dev_alloc(c, &c->tmp, count));
    // Predictor-side deposits (k_particles_predict_dep + k_tile_assemble + k_tile_fold_rho_chi).  Round 1: psi_middle source only, a wash on
    // config #2 (0.700 vs 0.696 ms/slice) and a gain on the big grids (config #3 4.50 vs 4.62, config #4 on one GPU 133.4 vs 150.4), so it was
    // chosen by size.  Round 2: the jx deposit and, in a slice's last pass, rho / susceptibility ride along (no tile kernel reads the particles
    // for them any more, no second stream in the B iterations): config #2 568 vs 600 us/slice, config #3 3.64 vs 3.79 ms, config #4 98.1 vs
    // 106.2 ms -- on by default for every size; QPIC_PREDICT_DEPOSIT=0 selects the tile kernels.
    if (!c->prm.force_general_deposit && c->jyjz_lists) {
        c->predict_deposit = c->predict_deposit_mode != 0;
    }
    {
        // rho / susceptibility from the last pass: the tile kernel it replaces runs on the second stream next to the push, so at 4 particles per
        // cell nothing is gained (config #2 564 vs 562 us/slice, config #3 3589 vs 3595); at 16 per cell the pass over the particles it saves
        // counts (config #4 on one GPU 96.0 vs 97.6 ms/slice).  Automatic: >= 8 particles per cell.  (Slab-decomposed sweeps: density of the
        // particles this rank was handed, which sit on 1 / world of the cells.)
        const size_t held = c->dist.on ? (size_t)c->dist.live0 * (size_t)c->prm.dist_world : count;
        c->fold_rc = c->fold_rc_mode == 1 || (c->fold_rc_mode == -1 && held >= 8 * (size_t)c->g.n2);
        // the same density decides how the predictor's per-cell sums leave the warp (QPIC_DENSE_CELLS = 0 / 1 forces it)
        const char * e_dense = getenv("QPIC_DENSE_CELLS");
        c->dense_cells = (e_dense && (e_dense[0] == '0' || e_dense[0] == '1')) ? e_dense[0] == '1' : held >= 8 * (size_t)c->g.n2;
    }
    if (c->predict_deposit && count > 0) {
        // experimental predictor-side deposit: per-cell corner sums, per-warp pieces, double-buffered mover counters, unsorted lists
        const size_t n2 = (size_t)c->g.n2, nwarps = (count + 31) / 32;
        if (!c->stay.corn) {
            QPIC_TRY(dev_alloc(c, &c->stay.corn, 4 * n2));
            QPIC_TRY(dev_alloc(c, &c->stay_cnt[0], n2));
            QPIC_TRY(dev_alloc(c, &c->stay_cnt[1], n2));
            QPIC_TRY(dev_alloc(c, &c->stay.lst, (size_t)BIN_K * n2));
            if (c->fold_jx) QPIC_TRY(dev_alloc(c, &c->stay.xcorn, 4 * n2));
            if (c->fold_jx && c->fold_rc) QPIC_TRY(dev_alloc(c, &c->stay.rcorn, 8 * n2));
        }
        QPIC_TRY(dev_alloc(c, &c->stay.lead, 4 * nwarps));
        QPIC_TRY(dev_alloc(c, &c->stay.trail, 4 * nwarps));
        if (c->fold_jx) {
            QPIC_TRY(dev_alloc(c, &c->stay.xlead, 4 * nwarps));
            QPIC_TRY(dev_alloc(c, &c->stay.xtrail, 4 * nwarps));
        }
        if (c->fold_jx && c->fold_rc) {
            QPIC_TRY(dev_alloc(c, &c->stay.rlead, 8 * nwarps));
            QPIC_TRY(dev_alloc(c, &c->stay.rtrail, 8 * nwarps));
        }
        c->stay.sign = -1.0;
        c->stay.allow_far = c->far_movers ? 1 : 0;
        far_rows(c, &c->stay.far_row_lo, &c->stay.far_row_n);
    }
    return QPIC_OK;
}

// Reads the device control block (synchronises).  Handles the two asynchronous events of a sweep:
//   stop condition  -> the sweep is over (src/System_3d.cpp:588-591);
//   fast-path abort -> roll the slice back (By/Bz of the slice start = row i+1 of the 3-D arrays, or zero for the first
//                      slice; nothing else the slice wrote is read before being rewritten) and schedule the general path.
int handle_ctl(qpic_ctx * c, const Ctl & seen) {
    if (c->dist.on) {
        static const char * why[] = {"", "an exchange barrier timed out (a peer stopped or the ranks diverged)", "more emigrants in one slice than the inbox holds",
                                     "particle_capacity exceeded after migration", "a particle moved further than the neighbouring slab"};
        const int e = seen.dist_error;
        if (e != 0) return fail(c, QPIC_ERR_STATE, "distributed sweep failed on rank %d: %s", c->dist.v.rank, why[e > 0 && e < 5 ? e : 0]);
        if (seen.fast_fail != 0) {
            return fail(c, QPIC_ERR_UNSUPPORTED,
                        "rank %d: a half step carried a particle beyond the %d halo rows of its slab (or a far-mover list overflowed); the sorted-deposit fallback "
                        "is not available in distributed sweeps",
                        c->dist.v.rank, DIST_HALO);
        }
    }
    if (seen.stop_tag != 0) {
        c->stopped_at = (seen.stop_tag - 1) / STOP_TAG_STRIDE;
        c->cur = c->stopped_at;
        c->done = true;
        return QPIC_OK;
    }
    if (seen.fast_fail != 0) {
        // (a lagged snapshot is as good as a fresh one here: once the flag is up every later kernel, including the slice counter, is
        // a no-op until the host clears it)
        const int slice = seen.slice;
        const size_t n2 = (size_t)c->g.n2;
        if (slice < c->g.nx - 1) {
            QPIC_CUDA(c, cudaMemcpyAsync(c->f2[QPIC_F2_BY_CUR], c->f3[QPIC_F3_BY] + n2 * (slice + 1), sizeof(double) * n2, cudaMemcpyDeviceToDevice, c->stream));
            QPIC_CUDA(c, cudaMemcpyAsync(c->f2[QPIC_F2_BZ_CUR], c->f3[QPIC_F3_BZ] + n2 * (slice + 1), sizeof(double) * n2, cudaMemcpyDeviceToDevice, c->stream));
        } else {
            QPIC_CUDA(c, cudaMemsetAsync(c->f2[QPIC_F2_BY_CUR], 0, sizeof(double) * n2, c->stream));
            QPIC_CUDA(c, cudaMemsetAsync(c->f2[QPIC_F2_BZ_CUR], 0, sizeof(double) * n2, c->stream));
        }
        // (the aborted slice may have left far-mover entries behind: its deposits, which clear the counters, became no-ops)
        for (int * cnt : c->far_cnt) {
            if (cnt) QPIC_CUDA(c, cudaMemsetAsync(cnt, 0, sizeof(int) * n2, c->stream));
        }
        for (int * cnt : c->stay_cnt) {
            if (cnt) QPIC_CUDA(c, cudaMemsetAsync(cnt, 0, sizeof(int) * n2, c->stream));
        }
        *c->ctl_host = seen;
        c->ctl_host->fast_fail = 0;
        QPIC_CUDA(c, cudaMemcpyAsync(c->ctl, c->ctl_host, sizeof(Ctl), cudaMemcpyHostToDevice, c->stream));
        QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
        c->cur = slice;
        c->done = false;
        c->general_left = c->general_penalty;
        if (c->general_penalty < 4096) c->general_penalty *= 2;
        c->fast_aborts++;
    }
    return QPIC_OK;
}

int poll_ctl(qpic_ctx * c) {
    c->snap_pending = false;  // superseded by the fresh copy below
    QPIC_CUDA(c, cudaMemcpyAsync(c->ctl_host, c->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->since_poll = 0;
    const Ctl seen = *c->ctl_host;
    return handle_ctl(c, seen);
}

// the non-blocking variant: act on the snapshot requested one interval ago, request the next one
int poll_ctl_lagged(qpic_ctx * c) {
    c->since_poll = 0;
    if (c->snap_pending) {
        QPIC_CUDA(c, cudaEventSynchronize(c->ev_snap));
        c->snap_pending = false;
        const Ctl seen = *c->ctl_snap;
        QPIC_TRY(handle_ctl(c, seen));
        if (c->done || seen.fast_fail != 0) return QPIC_OK;  // (after a rollback the stream is idle: nothing to snapshot)
    }
    // The copy runs on its own stream behind an event of the sweep's stream: a transfer queued on the compute stream itself sits between two
    // graph launches and costs the sweep ~1.3 ms per poll (21 us per slice at one poll per 64 slices, gpurun_out/r87: 588 vs 567 us/slice).
    // The flags it reads are sticky and the slice counter stops with them, so a snapshot taken a few kernels later is as good.
    QPIC_CUDA(c, cudaEventRecord(c->ev_poll_pos, c->stream));
    QPIC_CUDA(c, cudaStreamWaitEvent(c->poll_stream, c->ev_poll_pos, 0));
    QPIC_CUDA(c, cudaMemcpyAsync(c->ctl_snap, c->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, c->poll_stream));
    QPIC_CUDA(c, cudaEventRecord(c->ev_snap, c->poll_stream));
    c->snap_pending = true;
    return QPIC_OK;
}

int build_graph(qpic_ctx * c, bool fast) {
#ifdef QPIC_EMU
    (void)c;
    (void)fast;
    return QPIC_OK;
#else
    cudaGraphExec_t & exec = fast ? c->graph_exec_fast : c->graph_exec;
    if (exec) return QPIC_OK;
    QPIC_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const long before = c->launch_counter;
    int rc = enqueue_slice_full(c, fast);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (rc != QPIC_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (e != cudaSuccess) return fail(c, QPIC_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    (fast ? c->launches_per_slice_fast : c->launches_per_slice) = c->launch_counter - before;
    c->launch_counter = before;
    (fast ? c->graph_fast : c->graph) = graph;
    QPIC_CUDA(c, cudaGraphInstantiate(&exec, graph, 0));
    return QPIC_OK;
#endif
}

// queues up to `count` slices starting at the host's idea of the current slice; polls the device every
// stop_check_interval slices (a poll may move c->cur back up after a fast-path abort)
int launch_slices(qpic_ctx * c, int count) {
    const int M = c->prm.magnetic_field_iterations;
    int launched = 0;
    while (!c->done && launched < count) {
        const bool fast = c->fast_capable && c->general_left == 0;
        if (c->cur == 0) {
            // slice 0: the loop body up to the `i == 0` break (src/System_3d.cpp:334-411)
            QPIC_TRY(enqueue_slice_head(c, fast, false, true));
            c->done = true;
            launched++;
            break;
        }
#ifndef QPIC_EMU
        if (c->prm.use_graph) {
            QPIC_TRY(build_graph(c, fast));
            QPIC_CUDA(c, cudaGraphLaunch(fast ? c->graph_exec_fast : c->graph_exec, c->stream));
            c->launch_counter += fast ? c->launches_per_slice_fast : c->launches_per_slice;
        } else
#endif
        {
            QPIC_TRY(enqueue_slice_full(c, fast));
        }
        if (!fast && c->general_left > 0) c->general_left--;
        c->cur--;
        launched++;
        if (M > 0 && ++c->since_poll >= c->prm.stop_check_interval) QPIC_TRY(c->lagged ? poll_ctl_lagged(c) : poll_ctl(c));
    }
    return QPIC_OK;
}

// brings the device to the position the caller has asked for (c->target), resolving pending fast-path aborts
int resolve(qpic_ctx * c) {
    const int M = c->prm.magnetic_field_iterations;
    for (int guard = 0; guard < 1 << 20; guard++) {
        if (M > 0) QPIC_TRY(poll_ctl(c));
        else QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->stopped_at >= 0) return QPIC_OK;
        const bool head_pending = (c->target < 0) && !c->done;  // target -1: slice 0 processed as well
        if (c->cur <= (c->target < 0 ? 0 : c->target) && !head_pending) return QPIC_OK;
        QPIC_TRY(launch_slices(c, c->cur - (c->target < 0 ? -1 : c->target)));
    }
    return fail(c, QPIC_ERR_STATE, "resolve did not converge");
}

}  // namespace

// ================================================================================================ C ABI

extern "C" {

int qpic_abi_version(void) { return QPIC_ABI_VERSION; }

#ifndef QPIC_SOURCE_HASH
#define QPIC_SOURCE_HASH "unknown"
#endif
const char * qpic_source_hash(void) {
    static const char tag[] = QPIC_SOURCE_HASH;  // "QPICSRC:<16 hex digits>" (the prefix lets build() find it in the file)
    return std::strncmp(tag, "QPICSRC:", 8) == 0 ? tag + 8 : tag;
}

void qpic_default_params(qpic_params * p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->psi_threshold = 1e-4;
    p->magnetic_field_D = 10.0;
    p->threshold_B = 100.0;
    p->device = 0;
    p->keep_slices_3d = 0;
    p->use_graph = 1;
    p->stop_check_interval = 64;
    p->dist_rank = 0;
    p->dist_world = 1;
    p->particle_capacity = 0;
    p->keep_xy_cuts = 0;
    p->xy_z0 = -1.0;
}

const char * qpic_last_error(const qpic_ctx * ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void qpic_destroy(qpic_ctx * c) {
    if (!c) return;
    cudaSetDevice(c->prm.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
#ifndef QPIC_EMU
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->graph) cudaGraphDestroy(c->graph);
    if (c->graph_exec_fast) cudaGraphExecDestroy(c->graph_exec_fast);
    if (c->graph_fast) cudaGraphDestroy(c->graph_fast);
#endif
    for (void * p : c->allocations) cudaFree(p);
#ifndef QPIC_EMU
    // The predictor stores its half-step outputs with an L2 evict-last policy (QPIC_L2_KEEP): those lines keep their priority after the arrays
    // are gone; hand them back so that a later context of the process starts with the whole L2.
    cudaCtxResetPersistingL2Cache();
    for (void * m : c->dist.opened) cudaIpcCloseMemHandle(m);
#endif
    if (c->ctl_host) cudaFreeHost(c->ctl_host);
    if (c->ctl_snap) cudaFreeHost(c->ctl_snap);
    if (c->ev_snap) cudaEventDestroy(c->ev_snap);
    if (c->ev_poll_pos) cudaEventDestroy(c->ev_poll_pos);
    if (c->poll_stream) {
        cudaStreamSynchronize(c->poll_stream);
        cudaStreamDestroy(c->poll_stream);
    }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamDestroy(c->copy_stream);
    }
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int create_impl(qpic_ctx * c) {
    const qpic_params & p = c->prm;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        return fail(c, QPIC_ERR_CUDA, "no CUDA device available: qpic_b200 has no CPU fallback");
    }
    if (p.device < 0 || p.device >= ndev) return fail(c, QPIC_ERR_INVALID, "device %d out of range (%d devices)", p.device, ndev);
    QPIC_CUDA(c, cudaSetDevice(p.device));
    QPIC_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    QPIC_CUDA(c, cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    QPIC_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    QPIC_CUDA(c, cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
    for (int e = 0; e < 64; e++) {
        cudaEvent_t ev;
        QPIC_CUDA(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        c->ev_pool.push_back(ev);
    }
    QPIC_CUDA(c, cudaEventCreate(&c->ev0));
    QPIC_CUDA(c, cudaEventCreate(&c->ev1));

    Geom & g = c->g;
    g.nx = p.nx;
    g.ny = p.ny;
    g.nz = p.nz;
    g.n2 = p.ny * p.nz;
    g.dx = p.lx / p.nx;  // src/System_3d.cpp:93-95
    g.dy = p.ly / p.ny;
    g.dz = p.lz / p.nz;
    g.ly = p.ly;
    g.lz = p.lz;
    g.inv_dx = 1.0 / g.dx;
    g.inv_dy = 1.0 / g.dy;
    g.inv_dz = 1.0 / g.dz;
    {
        int e;
        g.exact_dy = std::frexp(g.dy, &e) == 0.5 ? 1 : 0;
        g.exact_dz = std::frexp(g.dz, &e) == 0.5 ? 1 : 0;
    }

    const size_t n2 = (size_t)g.n2, n3 = n2 * g.nx;
    qpic_ctx::Dist & D = c->dist;
    D.on = p.dist_world > 1;
    if (D.on) {
        const int G = p.dist_world;
        const bool pow2 = (g.ny & (g.ny - 1)) == 0 && (g.nz & (g.nz - 1)) == 0 && g.ny >= 64 && g.nz >= 64;
        if (G > DIST_MAX || p.dist_rank < 0 || p.dist_rank >= G) return fail(c, QPIC_ERR_INVALID, "dist_rank / dist_world out of range (at most %d ranks)", DIST_MAX);
        if (!pow2 || g.ny % G != 0 || (g.ny / G) % TILE != 0 || g.ny / G < 2 * TILE || g.nz % G != 0 || p.force_general_deposit || p.keep_slices_3d) {
            return fail(c, QPIC_ERR_UNSUPPORTED,
                        "distributed sweeps need power-of-two ny, nz >= 64, ny / world a multiple of %d and >= %d, the sort-free deposit and no per-slice 3-D copies",
                        TILE, 2 * TILE);
        }
        DistView & v = D.v;
        v.rank = p.dist_rank;
        v.world = G;
        v.below = (p.dist_rank + G - 1) % G;
        v.above = (p.dist_rank + 1) % G;
        v.ny_loc = g.ny / G;
        v.row0 = p.dist_rank * v.ny_loc;
        const int nzh = g.nz / 2 + 1;
        D.cpr_half = (nzh + G - 1) / G;
        D.cpr_full = g.nz / G;
        D.col0_half = p.dist_rank * D.cpr_half;
        D.ncols_half = nzh - D.col0_half < 0 ? 0 : (nzh - D.col0_half < D.cpr_half ? nzh - D.col0_half : D.cpr_half);
        D.col0_full = p.dist_rank * D.cpr_full;
        D.ncols_full = D.cpr_full;
    }
    QPIC_TRY(make_fft(c, g.ny, &c->fy));
    QPIC_TRY(make_fft(c, g.nz, &c->fz));
    QPIC_TRY(make_symbol_tables(c));

    // launch geometry of the transforms
    {
        const long long prod = (long long)g.ny * g.nz;
        c->inv_scale = ((prod & (prod - 1)) == 0) ? 1.0 / (double)prod : 0.0;
    }
    c->pow2_y = c->fy.plan.log2n > 0;
    c->pow2_z = c->fz.plan.log2n > 0;
    KernelLimits lim_real[3], lim_cplx[3], lim_cols[3];
#define P2_LIMITS(M)                                                                                                                                  \
    lim_real[M] = kernel_limits(k_p2_rows_r2c<false, M>, k_p2_rows_r2c<true, M>, k_p2_rows_c2r<false, M>, k_p2_rows_c2r<true, M>);                   \
    lim_cplx[M] = kernel_limits(k_p2_rows_b_forward<false, false, M>, k_p2_rows_b_forward<false, true, M>, k_p2_rows_b_forward<true, false, M>,      \
                                k_p2_rows_b_backward<false, M>, k_p2_rows_b_backward<true, M>, k_p2_rows_c2c<M>);                                    \
    lim_cols[M] = kernel_limits(k_p2_cols_solve<M>);
    P2_LIMITS(0)
    P2_LIMITS(1)
    P2_LIMITS(2)
#undef P2_LIMITS
    QPIC_TRY(size_fft_launch(c, c->fz.plan, (row_count(c) + 1) / 2, &c->rows_real, "row transform", lim_real));
    QPIC_TRY(size_fft_launch(c, c->fz.plan, row_count(c), &c->rows_cplx, "row transform", lim_cplx));
    QPIC_TRY(size_fft_launch(c, c->fy.plan, D.on ? (D.ncols_half > 0 ? D.ncols_half : 1) : g.nz / 2 + 1, &c->cols_half, "column transform", lim_cols));
    QPIC_TRY(size_fft_launch(c, c->fy.plan, D.on ? D.ncols_full : g.nz, &c->cols_full, "column transform", lim_cols));
    if (!c->pow2_y) {
        // generic engine: two adjacent columns per block when they fit, so that global accesses are 32-byte sectors
        const int cpb = (sizeof(double2) * 2 * 2 * (size_t)g.ny <= SMEM_DEFAULT_LIMIT) ? 2 : 1;
        for (qpic_ctx::FftLaunch * L : {&c->cols_half, &c->cols_full}) {
            const int ncols = (L == &c->cols_half) ? g.nz / 2 + 1 : g.nz;
            L->spb = cpb;
            L->threads = pick_threads(cpb * g.ny / 4);
            L->smem = sizeof(double2) * 2 * (size_t)cpb * g.ny;
            L->blocks = (ncols + cpb - 1) / cpb;
        }
    }
    {
        size_t rows_max = c->rows_real.smem > c->rows_cplx.smem ? c->rows_real.smem : c->rows_cplx.smem;
        size_t cols_max = c->cols_half.smem > c->cols_full.smem ? c->cols_half.smem : c->cols_full.smem;
        if (rows_max > SMEM_DEFAULT_LIMIT) {
            const int b = (int)rows_max;
            QPIC_CUDA(c, cudaFuncSetAttribute(k_fft_rows_r2c<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
            QPIC_CUDA(c, cudaFuncSetAttribute(k_fft_rows_c2r<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
            QPIC_CUDA(c, cudaFuncSetAttribute(k_fft_rows_b_forward<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
            QPIC_CUDA(c, cudaFuncSetAttribute(k_fft_rows_b_backward<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
            QPIC_CUDA(c, cudaFuncSetAttribute(k_fft_rows_c2c<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
#define P2_SMEM(M)                                                                                                            \
    QPIC_CUDA(c, cudaFuncSetAttribute((k_p2_rows_r2c<false, M>), cudaFuncAttributeMaxDynamicSharedMemorySize, b));              \
    QPIC_CUDA(c, cudaFuncSetAttribute((k_p2_rows_r2c<true, M>), cudaFuncAttributeMaxDynamicSharedMemorySize, b));               \
    QPIC_CUDA(c, cudaFuncSetAttribute((k_p2_rows_c2r<false, M>), cudaFuncAttributeMaxDynamicSharedMemorySize, b));              \
    QPIC_CUDA(c, cudaFuncSetAttribute((k_p2_rows_c2r<true, M>), cudaFuncAttributeMaxDynamicSharedMemorySize, b));               \
    QPIC_CUDA(c, cudaFuncSetAttribute((k_p2_rows_b_forward<false, false, M>), cudaFuncAttributeMaxDynamicSharedMemorySize, b)); \
    QPIC_CUDA(c, cudaFuncSetAttribute((k_p2_rows_b_forward<false, true, M>), cudaFuncAttributeMaxDynamicSharedMemorySize, b));  \
    QPIC_CUDA(c, cudaFuncSetAttribute((k_p2_rows_b_forward<true, false, M>), cudaFuncAttributeMaxDynamicSharedMemorySize, b));  \
    QPIC_CUDA(c, cudaFuncSetAttribute((k_p2_rows_b_backward<false, M>), cudaFuncAttributeMaxDynamicSharedMemorySize, b));       \
    QPIC_CUDA(c, cudaFuncSetAttribute((k_p2_rows_b_backward<true, M>), cudaFuncAttributeMaxDynamicSharedMemorySize, b));        \
    QPIC_CUDA(c, cudaFuncSetAttribute(k_p2_rows_c2c<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, b));
            P2_SMEM(0)
            P2_SMEM(1)
            P2_SMEM(2)
#undef P2_SMEM
        }
        if (cols_max > SMEM_DEFAULT_LIMIT) {
            QPIC_CUDA(c, cudaFuncSetAttribute(k_fft_cols_solve<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cols_max));
            QPIC_CUDA(c, cudaFuncSetAttribute(k_p2_cols_solve<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cols_max));
            QPIC_CUDA(c, cudaFuncSetAttribute(k_p2_cols_solve<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cols_max));
            QPIC_CUDA(c, cudaFuncSetAttribute(k_p2_cols_solve<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cols_max));
        }
    }
    QPIC_CUDA(c, cudaFuncSetAttribute(k_tile_cell_deposit<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 12 * LIST_BY * 34)));
    // the inverse-symbol tables of the two operators of the sweep (psi: D = 0, half spectrum; By + i Bz: D, full spectrum)
    if (c->pow2_y) {
        const double * unused = nullptr;
        QPIC_TRY(get_inv_symbol(c, 0.0, false, &unused));
        QPIC_TRY(get_inv_symbol(c, p.magnetic_field_D, true, &unused));
    }

    QPIC_TRY(dev_alloc(c, &c->ctl, 1));
    QPIC_TRY(dev_alloc(c, &c->ctl_scratch, 1));
    QPIC_CUDA(c, cudaMallocHost(&c->ctl_host, sizeof(Ctl)));
    QPIC_CUDA(c, cudaMallocHost(&c->ctl_snap, sizeof(Ctl)));
    QPIC_CUDA(c, cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming));
    QPIC_CUDA(c, cudaEventCreateWithFlags(&c->ev_poll_pos, cudaEventDisableTiming));
    QPIC_CUDA(c, cudaStreamCreateWithFlags(&c->poll_stream, cudaStreamNonBlocking));
    if (D.on) {
        // the exchange region: same layout on every rank
        DistView & v = D.v;
        DistLayout & L = v.lay;
        v.mig_cap = 64 * g.nz;  // particles that may cross one slab edge in one slice
        size_t off = 0;
        auto take = [&off](size_t bytes) {
            const size_t at = off;
            off += (bytes + 255) / 256 * 256;
            return at;
        };
        L.flags = take(sizeof(long long) * DIST_MAX);
        L.stops = take(sizeof(int) * 2 * DIST_MAX);
        L.partials = take(sizeof(double) * 4 * DIST_MAX);
        L.inbox_count = take(sizeof(int) * 2);
        const int cpr_max = D.cpr_half > D.cpr_full ? D.cpr_half : D.cpr_full;
        L.cols = take(sizeof(double2) * (size_t)g.ny * cpr_max);
        L.rbuf = take(sizeof(double) * RB_COUNT * 2 * DIST_HALO * g.nz);
        L.inbox = take(sizeof(double) * 2 * 6 * (size_t)v.mig_cap);
        for (int f = 0; f < DF_COUNT; f++) L.field[f] = take(sizeof(double) * n2);
        L.total = off;
        QPIC_TRY(dev_alloc(c, &D.region, L.total));
        for (int q = 0; q < DIST_MAX; q++) v.base[q] = nullptr;
        v.base[v.rank] = D.region;
        QPIC_TRY(dev_alloc(c, &D.out_count, 2));
        QPIC_TRY(dev_alloc(c, &D.epoch, 1));
        v.epoch = D.epoch;
        v.error = &c->ctl->dist_error;
        const int in_region[DF_COUNT] = {QPIC_F2_PSI, QPIC_F2_PSI_MIDDLE, QPIC_F2_BY_CUR, QPIC_F2_BZ_CUR, QPIC_F2_JX};
        for (int f = 0; f < DF_COUNT; f++) c->f2[in_region[f]] = reinterpret_cast<double *>(D.region + L.field[f]);
    }
    for (int f = 0; f < QPIC_F2_COUNT; f++) {
        if (!c->f2[f]) QPIC_TRY(dev_alloc(c, &c->f2[f], n2));
    }
    for (int f = 0; f < QPIC_F3_COUNT; f++) {
        const bool bunch = (f == QPIC_F3_RHO_BUNCH || f == QPIC_F3_JX_BUNCH);  // allocated on upload
        const bool slices = (f >= QPIC_F3_PSI);
        if (bunch || (slices && !p.keep_slices_3d)) continue;
        QPIC_TRY(dev_alloc(c, &c->f3[f], n3));
    }
    {
        c->jyjz_lists = true;
        // QPIC_PREDICT_DEPOSIT = 0 selects the tile kernels (k_tile_stay_deposit + jx tile deposit) instead of the predictor-side deposits
        const char * e4 = getenv("QPIC_PREDICT_DEPOSIT");
        c->predict_deposit_mode = (e4 && (e4[0] == '0' || e4[0] == '1')) ? e4[0] - '0' : -1;
        c->predict_deposit = false;
        const char * e7 = getenv("QPIC_FOLD_RHO_CHI");
        c->fold_rc_mode = (e7 && (e7[0] == '0' || e7[0] == '1')) ? e7[0] - '0' : -1;
        const char * e6 = getenv("QPIC_FOLD_JX");
        c->fold_jx = !(e6 && e6[0] == '0');
        const char * e5 = getenv("QPIC_FUSE_GAMMA");
        c->fuse_gamma = !(e5 && e5[0] == '0');
        const char * e3 = getenv("QPIC_TRACK_GAMMA_PX");
        c->track_gamma_px = e3 && e3[0] == '1';
        const char * e9 = getenv("QPIC_FAR_MOVERS");
        c->far_movers = !(e9 && e9[0] == '0') && !p.force_general_deposit;
        if (c->far_movers) {
            QPIC_TRY(dev_alloc(c, &c->far_cnt[0], n2));
            QPIC_TRY(dev_alloc(c, &c->far_cnt[1], n2));
            QPIC_TRY(dev_alloc(c, &c->far_lst, (size_t)FAR_K * n2));
        }
        if (c->jyjz_lists && !p.force_general_deposit) {
            QPIC_TRY(dev_alloc(c, &c->bin_cnt, n2));
            QPIC_TRY(dev_alloc(c, &c->bin_lst, (size_t)BIN_K * n2));
            // (distributed sweeps reduce jy_next / jz_next over the slab edges first, so their row kernel builds the sources)
            if (!D.on && c->pow2_z) QPIC_TRY(dev_alloc(c, &c->b_src, n2));
        } else {
            c->jyjz_lists = false;
        }
    }
    if (p.keep_xy_cuts) {
        for (int f = 0; f < 5; f++) QPIC_TRY(dev_alloc(c, &c->xy[f], (size_t)g.nx * g.ny));
    }
    QPIC_TRY(dev_alloc(c, &c->fft_in, n2));
    QPIC_TRY(dev_alloc(c, &c->fft_in2, n2));
    QPIC_TRY(dev_alloc(c, &c->raw, n2));
    QPIC_TRY(dev_alloc(c, &c->spec, n2));
    QPIC_TRY(dev_alloc(c, &c->spec_row0, (size_t)g.nz / 2 + 1));
    QPIC_TRY(dev_alloc(c, &c->spec_colsum, (size_t)g.nz / 2 + 1));
    QPIC_TRY(dev_alloc(c, &c->cols_ticket, 1));
    QPIC_TRY(dev_alloc(c, &c->eos_ticket, 1));
    QPIC_TRY(dev_alloc(c, &c->partials, 4 * DIST_MAX));
    c->nchunks = blocks_for(n2 + 1, SCAN_CHUNK);  // + the "emigrated" bucket of distributed sweeps
    QPIC_TRY(dev_alloc(c, &c->count, n2 + 1));
    QPIC_TRY(dev_alloc(c, &c->cursor, n2 + 1));
    QPIC_TRY(dev_alloc(c, &c->sums, (size_t)c->nchunks));
    QPIC_TRY(dev_alloc(c, &c->start_b, n2 + 2));
    QPIC_TRY(dev_alloc(c, &c->start_m, n2 + 2));
    QPIC_TRY(alloc_particles(c, 0));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->cur = g.nx - 1;
    c->target = g.nx - 1;
    // the tile deposit needs unambiguous minimal-image cell offsets, and its halo cells (up to index TILE of a grid that fits in
    // one tile) have to stay below 2 n for wrap_near: n >= 9
    c->fast_capable = !p.force_general_deposit && g.ny >= 9 && g.nz >= 9;
    if (D.on && !(c->pow2_y && c->pow2_z)) return fail(c, QPIC_ERR_UNSUPPORTED, "distributed sweeps need the power-of-two transform engine");
#ifndef QPIC_EMU
    if (D.on) {
        // CUDA loads kernels lazily, and loading one synchronises the context: with several ranks in ONE process (shared
        // context) that would wait for ever on a peer's spinning barrier kernel.  Load everything the sweep launches now.
        cudaFuncAttributes fa;
#define QPIC_PRELOAD(k) QPIC_CUDA(c, cudaFuncGetAttributes(&fa, k))
        QPIC_PRELOAD(k_xchg_barrier);
        QPIC_PRELOAD(k_halo_reduce_send);
        QPIC_PRELOAD(k_halo_reduce_add);
        QPIC_PRELOAD(k_halo_fill);
        QPIC_PRELOAD(k_migrate_publish);
        QPIC_PRELOAD(k_migrate_unpack);
        QPIC_PRELOAD((k_tile_deposit<1, false>));
        QPIC_PRELOAD((k_tile_deposit<2, false>));
        QPIC_PRELOAD((k_tile_deposit<3, false>));
        QPIC_PRELOAD((k_tile_deposit<2, true, true>));
        QPIC_PRELOAD(k_tile_stay_deposit);
        QPIC_PRELOAD(k_tile_cell_deposit<16>);
        QPIC_PRELOAD(k_tile_cell_deposit<32>);
        QPIC_PRELOAD(k_particles_predict_dep<false>);
        QPIC_PRELOAD(k_particles_predict_dep<true>);
        QPIC_PRELOAD((k_particles_predict_dep<false, 1>));
        QPIC_PRELOAD((k_particles_predict_dep<true, 1>));
        QPIC_PRELOAD((k_particles_predict_dep<false, 2>));
        QPIC_PRELOAD((k_particles_predict_dep<true, 2>));
        QPIC_PRELOAD((k_particles_predict_dep<false, 1, true>));
        QPIC_PRELOAD((k_particles_predict_dep<true, 1, true>));
        QPIC_PRELOAD((k_particles_predict_dep<false, 2, true>));
        QPIC_PRELOAD((k_particles_predict_dep<true, 2, true>));
        QPIC_PRELOAD(k_tile_fold_rho_chi);
        QPIC_PRELOAD(k_tile_assemble<false>);
        QPIC_PRELOAD(k_tile_assemble<true>);
        QPIC_PRELOAD((k_particles_predict<true, true>));
        QPIC_PRELOAD(k_particles_predict<true>);
        QPIC_PRELOAD(k_particles_gamma);
        QPIC_PRELOAD(k_particles_advance);
        QPIC_PRELOAD(k_rank_to_slots);
        QPIC_PRELOAD(k_locate_count);
        QPIC_PRELOAD(k_scan_chunk_sums);
        QPIC_PRELOAD(k_scan_sums);
        QPIC_PRELOAD(k_scan_apply);
        QPIC_PRELOAD(k_sort_fill);
        QPIC_PRELOAD(k_sort_rank);
        QPIC_PRELOAD(k_end_of_slice);
        QPIC_PRELOAD(k_init_a_sqr);
#define P2_PRELOAD(M)                                        \
    QPIC_PRELOAD((k_p2_rows_r2c<true, M>));                  \
    QPIC_PRELOAD((k_p2_rows_c2r<true, M>));                  \
    QPIC_PRELOAD(k_p2_cols_solve<M>);                        \
    QPIC_PRELOAD((k_p2_rows_b_forward<true, false, M>));     \
    QPIC_PRELOAD((k_p2_rows_b_backward<true, M>));
        P2_PRELOAD(0)
        P2_PRELOAD(1)
        P2_PRELOAD(2)
#undef P2_PRELOAD
#undef QPIC_PRELOAD
    }
#endif
    return QPIC_OK;
}

int qpic_create(const qpic_params * params, qpic_ctx ** out) {
    if (!out) return fail(nullptr, QPIC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!params) return fail(nullptr, QPIC_ERR_INVALID, "params is NULL");
    if (params->nx < 1 || params->ny < 1 || params->nz < 1) return fail(nullptr, QPIC_ERR_INVALID, "grid dimensions must be positive");
    if (!(params->lx > 0) || !(params->ly > 0) || !(params->lz > 0)) return fail(nullptr, QPIC_ERR_INVALID, "box sizes must be positive");
    if (params->magnetic_field_iterations < 0) {
        return fail(nullptr, QPIC_ERR_INVALID, "Magnetic field iterations should be non-negative");  // src/System_3d.cpp:77-79
    }
    if (params->magnetic_field_iterations >= STOP_TAG_STRIDE) return fail(nullptr, QPIC_ERR_INVALID, "too many magnetic field iterations");
    if ((double)params->ny * params->nz > 1.0e9 || (double)params->nx * params->ny * params->nz > 8.0e18) {
        return fail(nullptr, QPIC_ERR_INVALID, "grid too large");
    }
    qpic_ctx * c = new (std::nothrow) qpic_ctx();
    if (!c) return fail(nullptr, QPIC_ERR_NOMEM, "out of host memory");
    c->prm = *params;
    if (c->prm.stop_check_interval <= 0) c->prm.stop_check_interval = 64;
#ifdef QPIC_EMU
    c->prm.use_graph = 0;
#endif
    int rc = create_impl(c);
    if (rc != QPIC_OK) {
        g_create_error = c->err;
        qpic_destroy(c);
        return rc;
    }
    *out = c;
    return QPIC_OK;
}

int qpic_set_laser_gaussian(qpic_ctx * c, double amplitude, double wx, double wy, double wz, double x0, double y0, double z0) {
    if (!c) return QPIC_ERR_INVALID;
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    const size_t total = (size_t)c->g.nx * c->g.n2;
    int blocks = blocks_for(total, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    QPIC_LAUNCH_FLAT(k_init_a_sqr, blocks, 256, 0, c->stream, c->g, c->f3[QPIC_F3_ASQR], amplitude, wx, wy, wz, x0, y0, z0);
    QPIC_TRY(check_launch(c, "k_init_a_sqr"));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

int qpic_upload_a_sqr(qpic_ctx * c, const double * a_sqr) {
    if (!c || !a_sqr) return fail(c, QPIC_ERR_INVALID, "a_sqr is NULL");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    QPIC_CUDA(c, cudaMemcpyAsync(c->f3[QPIC_F3_ASQR], a_sqr, sizeof(double) * (size_t)c->g.nx * c->g.n2, cudaMemcpyHostToDevice, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

int qpic_upload_bunch_sources(qpic_ctx * c, const double * rho_bunch, const double * jx_bunch) {
    if (!c) return QPIC_ERR_INVALID;
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    const size_t n3 = (size_t)c->g.nx * c->g.n2;
    const double * src[2] = {rho_bunch, jx_bunch};
    const int idx[2] = {QPIC_F3_RHO_BUNCH, QPIC_F3_JX_BUNCH};
    if ((c->graph_exec || c->graph_exec_fast) && (!c->f3[idx[0]] || !c->f3[idx[1]]) && (rho_bunch || jx_bunch)) {
        return fail(c, QPIC_ERR_STATE, "bunch sources must be uploaded before the first qpic_step of a context");
    }
    for (int t = 0; t < 2; t++) {
        if (src[t]) {
            if (!c->f3[idx[t]]) QPIC_TRY(dev_alloc(c, &c->f3[idx[t]], n3, false));
            QPIC_CUDA(c, cudaMemcpyAsync(c->f3[idx[t]], src[t], sizeof(double) * n3, cudaMemcpyHostToDevice, c->stream));
        } else if (c->f3[idx[t]]) {
            QPIC_CUDA(c, cudaMemsetAsync(c->f3[idx[t]], 0, sizeof(double) * n3, c->stream));
        }
    }
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

int qpic_init_wake_particles(qpic_ctx * c, const double * y, const double * z, const double * n, size_t count) {
    if (!c) return QPIC_ERR_INVALID;
    if (count > 0 && (!y || !z || !n)) return fail(c, QPIC_ERR_INVALID, "particle arrays are NULL");
    if (count > (size_t)std::numeric_limits<int>::max() - 1024) return fail(c, QPIC_ERR_INVALID, "too many particles for 32-bit particle indices");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    // distributed: the arrays hold `capacity` entries, the live count is kept on the device and changes with migration
    size_t capacity = count;
    if (c->dist.on) {
        capacity = c->prm.particle_capacity ? c->prm.particle_capacity : count + count / 4 + 65536;
        if (capacity < count) return fail(c, QPIC_ERR_INVALID, "particle_capacity is smaller than the particle count");
        if (capacity > (size_t)std::numeric_limits<int>::max() / 4 - 1024) return fail(c, QPIC_ERR_INVALID, "too many particles per rank");
        c->dist.live0 = (int)count;
    }
    if ((c->graph_exec || c->graph_exec_fast) && (int)capacity != c->P.np) {
        return fail(c, QPIC_ERR_STATE, "the particle count cannot change after the first qpic_step of a context");
    }
    if ((int)capacity != c->P.np || !c->P.y) {
        // (re)allocation: dev_alloc releases the previous arrays
        QPIC_TRY(alloc_particles(c, capacity));
    } else {
        double * zero[] = {c->P.py, c->P.pz, c->P.gamma, c->P.px, c->P.gam, c->P.pxs};
        for (double * p : zero) QPIC_CUDA(c, cudaMemsetAsync(p, 0, sizeof(double) * capacity, c->stream));
        QPIC_CUDA(c, cudaMemsetAsync(c->P.pyzn, 0, sizeof(double2) * capacity, c->stream));
    }
    if (count > 0) {
        QPIC_CUDA(c, cudaMemcpyAsync(c->P.y, y, sizeof(double) * count, cudaMemcpyHostToDevice, c->stream));
        QPIC_CUDA(c, cudaMemcpyAsync(c->P.z, z, sizeof(double) * count, cudaMemcpyHostToDevice, c->stream));
        QPIC_CUDA(c, cudaMemcpyAsync(c->P.nw, n, sizeof(double) * count, cudaMemcpyHostToDevice, c->stream));
    }
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->lists_dirty = true;
    c->began = false;
    return QPIC_OK;
}

int qpic_begin(qpic_ctx * c) {
    if (!c) return QPIC_ERR_INVALID;
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    const Geom & g = c->g;
    const size_t n2 = (size_t)g.n2, n3 = n2 * g.nx;
    // src/System_3d.cpp:265-316
    QPIC_CUDA(c, cudaMemsetAsync(c->f3[QPIC_F3_EX] + n2 * (size_t)(g.nx - 1), 0, sizeof(double) * n2, c->stream));
    QPIC_CUDA(c, cudaMemsetAsync(c->f3[QPIC_F3_BY], 0, sizeof(double) * n3, c->stream));
    QPIC_CUDA(c, cudaMemsetAsync(c->f3[QPIC_F3_BZ], 0, sizeof(double) * n3, c->stream));
    const int zero2[] = {QPIC_F2_PSI, QPIC_F2_PSI_PREV, QPIC_F2_PSI_MIDDLE, QPIC_F2_JY, QPIC_F2_JZ, QPIC_F2_JY_NEXT, QPIC_F2_JZ_NEXT,
                         QPIC_F2_BY_CUR, QPIC_F2_BZ_CUR, QPIC_F2_RHO_ION};
    for (int f : zero2) QPIC_CUDA(c, cudaMemsetAsync(c->f2[f], 0, sizeof(double) * n2, c->stream));
    for (int * cnt : c->stay_cnt) {
        if (cnt) QPIC_CUDA(c, cudaMemsetAsync(cnt, 0, sizeof(int) * n2, c->stream));
    }
    for (int * cnt : c->far_cnt) {
        if (cnt) QPIC_CUDA(c, cudaMemsetAsync(cnt, 0, sizeof(int) * n2, c->stream));
    }
    // The per-slice datasets psi, rho, jx, jy, jz live in the output file of ONE time iteration in the reference (Output_writer is
    // constructed per iteration, src/System_3d.cpp:209, and initialize_slice_array creates zero-filled datasets): slices a sweep
    // never reaches -- it stopped on |B| > threshold_B -- must read 0, not the previous sweep's values.
    for (int f = QPIC_F3_PSI; f < QPIC_F3_COUNT; f++) {
        if (c->f3[f]) QPIC_CUDA(c, cudaMemsetAsync(c->f3[f], 0, sizeof(double) * n3, c->stream));
    }
    for (double * cut : c->xy) {
        if (cut) QPIC_CUDA(c, cudaMemsetAsync(cut, 0, sizeof(double) * (size_t)g.nx * g.ny, c->stream));
    }
    Ctl init;
    init.slice = g.nx - 1;
    init.stop_tag = 0;
    init.fast_fail = 0;
    init.dist_error = 0;
    init.np1 = c->dist.on ? c->dist.live0 + 1 : 0;
    init.np_p1 = init.np1;
    init.mig_base1 = 0;
    init.pad = 0;
    if (c->dist.on && !c->dist.attached) return fail(c, QPIC_ERR_STATE, "distributed context: attach the peers' exchange regions first (qpic_dist_attach_*)");
    *c->ctl_host = init;
    QPIC_CUDA(c, cudaMemcpyAsync(c->ctl, c->ctl_host, sizeof(Ctl), cudaMemcpyHostToDevice, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    QPIC_TRY(enqueue_rebuild_base_lists(c));
    // rho_ion (src/System_3d.cpp:319-323)
    if (c->dist.on) {
        TileJob j = tile_job(c);
        j.nout = 1;
        j.o[0] = tile_output(c->P.nw_s, -1.0, 0, nullptr, c->f2[QPIC_F2_RHO_ION], 0);
        QPIC_TRY(launch_tile(c, c->stream, j, false, "k_tile_deposit_rho_ion"));
        QPIC_TRY(dist_reduce(c, c->stream, halo_job({c->f2[QPIC_F2_RHO_ION]}, {RB_RHO})));
    } else {
        QPIC_LAUNCH_FLAT(k_deposit_charge, blocks_for(g.n2, 128), 128, 0, c->stream, c->ctl_scratch, c->g, lists_base(c), c->P.nw_s, (const double *)nullptr,
                         (const double *)nullptr, (const double *)nullptr, 0, c->f2[QPIC_F2_RHO_ION]);
        QPIC_TRY(check_launch(c, "k_deposit_charge"));
    }
    c->cur = g.nx - 1;
    c->target = g.nx - 1;
    c->stopped_at = -1;
    c->done = false;
    c->finished = false;
    c->began = true;
    c->since_poll = 0;
    c->snap_pending = false;
    c->general_left = 0;
    c->general_penalty = 16;
    return QPIC_OK;
}

int qpic_step(qpic_ctx * c, int max_slices, int * done) {
    if (!c) return QPIC_ERR_INVALID;
    if (!c->began) return fail(c, QPIC_ERR_STATE, "qpic_step called before qpic_begin");
    if (max_slices < 0) return fail(c, QPIC_ERR_INVALID, "max_slices must be non-negative");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    if (c->lists_dirty) QPIC_TRY(enqueue_rebuild_base_lists(c));
    const long launches_before = c->launch_counter;
    QPIC_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    // the position the caller wants to reach: `target` counts down from nx-1; -1 means slice 0 has been processed too
    long want = (long)c->target - max_slices;
    if (want < -1) want = -1;
    c->target = (int)want;
    c->lagged = (done == nullptr);
    if (!c->done) QPIC_TRY(launch_slices(c, c->cur - c->target));
    c->lagged = false;
    if (done) {
        // the caller wants a definite answer: wait, and redo whatever a fast-path abort cancelled
        QPIC_TRY(resolve(c));
        *done = c->done ? 1 : 0;
    }
    QPIC_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    c->last_launches = c->launch_counter - launches_before;
    return QPIC_OK;
}

int qpic_finish(qpic_ctx * c) {
    if (!c) return QPIC_ERR_INVALID;
    if (!c->began) return fail(c, QPIC_ERR_STATE, "qpic_finish called before qpic_begin");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    QPIC_TRY(resolve(c));
    const size_t n2 = (size_t)c->g.n2;
    if (c->stopped_at >= 0) {
        // the reference wrote by(i), bz(i) of the stopping slice in place (src/System_3d.cpp:544, :572) and then broke out
        // of both loops (:582-591); slice 0 is still emitted (:622) with by(0) = bz(0) = 0.
        QPIC_CUDA(c, cudaMemcpyAsync(c->f3[QPIC_F3_BY] + n2 * c->stopped_at, c->f2[QPIC_F2_BY_CUR], sizeof(double) * n2, cudaMemcpyDeviceToDevice, c->stream));
        QPIC_CUDA(c, cudaMemcpyAsync(c->f3[QPIC_F3_BZ] + n2 * c->stopped_at, c->f2[QPIC_F2_BZ_CUR], sizeof(double) * n2, cudaMemcpyDeviceToDevice, c->stream));
        QPIC_TRY(enqueue_end_of_slice(c, 1, 0, 1));
    } else {
        QPIC_TRY(enqueue_end_of_slice(c, 0, 0, 0));
    }
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->finished = true;
    return QPIC_OK;
}

int qpic_solve_wakefield(qpic_ctx * c) {
    QPIC_TRY(qpic_begin(c));
    int done = 0;
    while (!done) QPIC_TRY(qpic_step(c, c->prm.stop_check_interval, &done));
    return qpic_finish(c);
}

int qpic_synchronize(qpic_ctx * c) {
    if (!c) return QPIC_ERR_INVALID;
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    QPIC_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    if (c->began && !c->finished) return resolve(c);
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

int qpic_current_slice(qpic_ctx * c) { return c ? c->cur : -1; }

int qpic_stopped_at(qpic_ctx * c) {
    if (!c) return -1;
    if (c->began && !c->finished) resolve(c);
    return c->stopped_at;
}

long qpic_fast_path_aborts(const qpic_ctx * c) { return c ? c->fast_aborts : 0; }

int qpic_last_step_stats(qpic_ctx * c, float * ms, long * kernel_launches) {
    if (!c) return QPIC_ERR_INVALID;
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    QPIC_CUDA(c, cudaEventSynchronize(c->ev1));
    QPIC_CUDA(c, cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
    if (ms) *ms = c->last_ms;
    if (kernel_launches) *kernel_launches = c->last_launches;
    return QPIC_OK;
}

// ---- state access

static int copy_out(qpic_ctx * c, const double * src, double * dst, size_t count) {
    if (!dst) return fail(c, QPIC_ERR_INVALID, "dst is NULL");
    if (!src) return fail(c, QPIC_ERR_STATE, "this array is not resident (see keep_slices_3d / qpic_upload_bunch_sources)");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    if (c->began && !c->finished) QPIC_TRY(resolve(c));
    QPIC_CUDA(c, cudaMemcpyAsync(dst, src, sizeof(double) * count, cudaMemcpyDeviceToHost, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

int qpic_download_field3d(qpic_ctx * c, qpic_field3d f, double * dst) {
    if (!c || f < 0 || f >= QPIC_F3_COUNT) return fail(c, QPIC_ERR_INVALID, "bad field id");
    const size_t n3 = (size_t)c->g.nx * c->g.n2;
    if ((f == QPIC_F3_RHO_BUNCH || f == QPIC_F3_JX_BUNCH) && !c->f3[f]) {
        if (!dst) return fail(c, QPIC_ERR_INVALID, "dst is NULL");
        std::memset(dst, 0, sizeof(double) * n3);  // never uploaded == zeros
        return QPIC_OK;
    }
    return copy_out(c, c->f3[f], dst, n3);
}

int qpic_download_slice3d(qpic_ctx * c, qpic_field3d f, int slice, double * dst) {
    if (!c || f < 0 || f >= QPIC_F3_COUNT) return fail(c, QPIC_ERR_INVALID, "bad field id");
    if (slice < 0 || slice >= c->g.nx) return fail(c, QPIC_ERR_INVALID, "slice out of range");
    const size_t n2 = (size_t)c->g.n2;
    if ((f == QPIC_F3_RHO_BUNCH || f == QPIC_F3_JX_BUNCH) && !c->f3[f]) {
        if (!dst) return fail(c, QPIC_ERR_INVALID, "dst is NULL");
        std::memset(dst, 0, sizeof(double) * n2);
        return QPIC_OK;
    }
    return copy_out(c, c->f3[f] ? c->f3[f] + n2 * slice : nullptr, dst, n2);
}

int qpic_download_field2d(qpic_ctx * c, qpic_field2d f, double * dst) {
    if (!c || f < 0 || f >= QPIC_F2_COUNT) return fail(c, QPIC_ERR_INVALID, "bad field id");
    return copy_out(c, c->f2[f], dst, (size_t)c->g.n2);
}

int qpic_upload_field2d(qpic_ctx * c, qpic_field2d f, const double * src) {
    if (!c || f < 0 || f >= QPIC_F2_COUNT || !src) return fail(c, QPIC_ERR_INVALID, "bad field id or NULL src");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    QPIC_CUDA(c, cudaMemcpyAsync(c->f2[f], src, sizeof(double) * (size_t)c->g.n2, cudaMemcpyHostToDevice, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

size_t qpic_particle_count(const qpic_ctx * c) {
    if (!c) return 0;
    if (c->dist.on) {
        // live particles of this rank right now (changes with migration); synchronises
        int np1 = 0;
        cudaSetDevice(c->prm.device);
        cudaStreamSynchronize(c->stream);
        if (cudaMemcpy(&np1, &c->ctl->np1, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
        return np1 > 0 ? (size_t)(np1 - 1) : (size_t)c->dist.live0;
    }
    return (size_t)c->P.np;
}

static double * particle_array(qpic_ctx * c, qpic_particle_field f) {
    switch (f) {
        case QPIC_P_Y: return c->P.y;
        case QPIC_P_Z: return c->P.z;
        case QPIC_P_PY: return c->P.py;
        case QPIC_P_PZ: return c->P.pz;
        case QPIC_P_N: return c->P.nw;
        case QPIC_P_GAMMA: return c->P.gamma;
        case QPIC_P_PX: return c->P.px;
        default: return nullptr;
    }
}

int qpic_download_particles(qpic_ctx * c, qpic_particle_field f, double * dst) {
    if (!c) return QPIC_ERR_INVALID;
    if (c->dist.on) return fail(c, QPIC_ERR_UNSUPPORTED, "particle arrays of a distributed sweep have no caller order");
    double * src = particle_array(c, f);
    if (!src) return fail(c, QPIC_ERR_INVALID, "bad particle field id");
    if (c->P.np == 0) return QPIC_OK;
    return copy_out(c, src, dst, (size_t)c->P.np);
}

int qpic_upload_particles(qpic_ctx * c, qpic_particle_field f, const double * src) {
    if (!c) return QPIC_ERR_INVALID;
    if (c->dist.on) return fail(c, QPIC_ERR_UNSUPPORTED, "particle arrays of a distributed sweep have no caller order");
    double * dst = particle_array(c, f);
    if (!dst || f == QPIC_P_GAMMA || f == QPIC_P_PX || !src) return fail(c, QPIC_ERR_INVALID, "only y, z, py, pz, n can be uploaded");
    if (c->P.np == 0) return QPIC_OK;
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    QPIC_CUDA(c, cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)c->P.np, cudaMemcpyHostToDevice, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    if (f == QPIC_P_Y || f == QPIC_P_Z) c->lists_dirty = true;
    return QPIC_OK;
}

int qpic_set_current_slice(qpic_ctx * c, int slice) {
    if (!c) return QPIC_ERR_INVALID;
    if (!c->began) return fail(c, QPIC_ERR_STATE, "qpic_set_current_slice called before qpic_begin");
    if (slice < 0 || slice >= c->g.nx) return fail(c, QPIC_ERR_INVALID, "slice out of range");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->ctl_host->slice = slice;
    c->ctl_host->stop_tag = 0;
    c->ctl_host->fast_fail = 0;
    QPIC_CUDA(c, cudaMemcpyAsync(c->ctl, c->ctl_host, sizeof(Ctl), cudaMemcpyHostToDevice, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->cur = slice;
    c->target = slice;
    c->done = false;
    c->finished = false;
    c->stopped_at = -1;
    return QPIC_OK;
}

// ---- building blocks

int qpic_solve_poisson_host(qpic_ctx * c, double * inout, double D) {
    if (!c || !inout) return fail(c, QPIC_ERR_INVALID, "inout is NULL");
    if (c->dist.on) return fail(c, QPIC_ERR_UNSUPPORTED, "host building blocks are single-GPU");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    const size_t n2 = (size_t)c->g.n2;
    QPIC_CUDA(c, cudaMemcpyAsync(c->fft_in, inout, sizeof(double) * n2, cudaMemcpyHostToDevice, c->stream));
    QPIC_TRY(enqueue_real_solve(c, c->stream, c->ctl_scratch, D));
    QPIC_CUDA(c, cudaMemcpyAsync(inout, c->raw, sizeof(double) * n2, cudaMemcpyDeviceToHost, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

int qpic_solve_helmholtz_pair_host(qpic_ctx * c, double * inout_a, double * inout_b, double D) {
    if (!c || !inout_a || !inout_b) return fail(c, QPIC_ERR_INVALID, "inout is NULL");
    if (c->dist.on) return fail(c, QPIC_ERR_UNSUPPORTED, "host building blocks are single-GPU");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    const size_t n2 = (size_t)c->g.n2;
    const int ny = c->g.ny, nz = c->g.nz;
    QPIC_CUDA(c, cudaMemcpyAsync(c->fft_in, inout_a, sizeof(double) * n2, cudaMemcpyHostToDevice, c->stream));
    QPIC_CUDA(c, cudaMemcpyAsync(c->fft_in2, inout_b, sizeof(double) * n2, cudaMemcpyHostToDevice, c->stream));
    {
        const qpic_ctx::FftLaunch & R = c->rows_cplx;
        if (c->pow2_z) {
            QPIC_LAUNCH(P2_PICK0(R.mode, k_p2_rows_c2c), R.blocks, R.threads, R.smem, c->stream, c->fz.plan.log2n, c->fz.tw, c->fft_in, c->fft_in2, c->spec, ny);
        } else {
            QPIC_LAUNCH(k_fft_rows_c2c<false>, R.blocks, R.threads, R.smem, c->stream, c->fz.plan, c->fz.tw, c->fft_in, c->fft_in2, c->spec, ny, nz, R.spb);
        }
        QPIC_TRY(check_launch(c, "k_fft_rows_c2c"));
    }
    QPIC_TRY(enqueue_complex_columns(c, c->stream, c->ctl_scratch, D));
    QPIC_TRY(enqueue_b_backward(c, c->stream, c->ctl_scratch, c->fft_in, c->fft_in2, std::numeric_limits<double>::infinity(), 0));
    QPIC_CUDA(c, cudaMemcpyAsync(inout_a, c->fft_in, sizeof(double) * n2, cudaMemcpyDeviceToHost, c->stream));
    QPIC_CUDA(c, cudaMemcpyAsync(inout_b, c->fft_in2, sizeof(double) * n2, cudaMemcpyDeviceToHost, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    // a NaN input would have tagged the scratch control block; clear it
    QPIC_CUDA(c, cudaMemsetAsync(c->ctl_scratch, 0, sizeof(Ctl), c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

// ---- slice-range transfers (streaming a_sqr in / fields out, used by the end-to-end path)

int qpic_upload_a_sqr_slices(qpic_ctx * c, int first_slice, int count, const double * src) {
    if (!c || !src) return fail(c, QPIC_ERR_INVALID, "src is NULL");
    if (first_slice < 0 || count < 0 || first_slice + count > c->g.nx) return fail(c, QPIC_ERR_INVALID, "slice range out of bounds");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    const size_t n2 = (size_t)c->g.n2;
    // on the copy stream, so that it overlaps slices already queued; slices queued AFTER this call wait for it
    QPIC_CUDA(c, cudaMemcpyAsync(c->f3[QPIC_F3_ASQR] + n2 * first_slice, src, sizeof(double) * n2 * count, cudaMemcpyHostToDevice, c->copy_stream));
    QPIC_CUDA(c, cudaEventRecord(c->ev_copy, c->copy_stream));
    QPIC_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
    return QPIC_OK;
}

int qpic_download_slices3d(qpic_ctx * c, qpic_field3d f, int first_slice, int count, double * dst) {
    if (!c || f < 0 || f >= QPIC_F3_COUNT || !dst) return fail(c, QPIC_ERR_INVALID, "bad field id or NULL dst");
    if (first_slice < 0 || count < 0 || first_slice + count > c->g.nx) return fail(c, QPIC_ERR_INVALID, "slice range out of bounds");
    if (!c->f3[f]) return fail(c, QPIC_ERR_STATE, "this array is not resident (see keep_slices_3d / qpic_upload_bunch_sources)");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    const size_t n2 = (size_t)c->g.n2;
    // waits for the slices queued so far, then runs on the copy stream while later slices compute
    QPIC_CUDA(c, cudaEventRecord(c->ev_copy, c->stream));
    QPIC_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_copy, 0));
    QPIC_CUDA(c, cudaMemcpyAsync(dst, c->f3[f] + n2 * first_slice, sizeof(double) * n2 * count, cudaMemcpyDeviceToHost, c->copy_stream));
    return QPIC_OK;
}

int qpic_download_xy_cut(qpic_ctx * c, qpic_field3d f, int first_slice, int count, double * dst) {
    if (!c || f < 0 || f >= QPIC_F3_COUNT || !dst) return fail(c, QPIC_ERR_INVALID, "bad field id or NULL dst");
    if (first_slice < 0 || count < 0 || first_slice + count > c->g.nx) return fail(c, QPIC_ERR_INVALID, "slice range out of bounds");
    if (count == 0) return QPIC_OK;
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    const size_t ny = (size_t)c->g.ny;
    // after the slices queued so far, on the copy stream
    QPIC_CUDA(c, cudaEventRecord(c->ev_copy, c->stream));
    QPIC_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_copy, 0));
    const double * src = nullptr;
    if (f >= QPIC_F3_PSI && c->xy[0]) {
        src = c->xy[f - QPIC_F3_PSI] + ny * first_slice;
    } else {
        if (!c->f3[f]) return fail(c, QPIC_ERR_STATE, "this array is not resident (see keep_xy_cuts / keep_slices_3d / qpic_upload_bunch_sources)");
        if (!c->xy_scratch) QPIC_TRY(dev_alloc(c, &c->xy_scratch, (size_t)c->g.nx * ny, false));
        const bool half = (f == QPIC_F3_BY || f == QPIC_F3_EZ || f == QPIC_F3_JZ);  // z-origin dz / 2 (src/System_3d.cpp:171-193)
        int blocks = blocks_for((size_t)count * ny, 256);
        if (blocks > 1184) blocks = 1184;
        QPIC_LAUNCH_FLAT(k_xy_cut, blocks, 256, 0, c->copy_stream, c->g, c->f3[f] + (size_t)c->g.n2 * first_slice, count, xy_cut(c, half ? 0.5 * c->g.dz : 0.0),
                         c->xy_scratch + ny * first_slice);
        c->launch_counter++;
        QPIC_CUDA(c, cudaGetLastError());
        src = c->xy_scratch + ny * first_slice;
    }
    QPIC_CUDA(c, cudaMemcpyAsync(dst, src, sizeof(double) * ny * count, cudaMemcpyDeviceToHost, c->copy_stream));
    return QPIC_OK;
}

int qpic_wait_transfers(qpic_ctx * c) {
    if (!c) return QPIC_ERR_INVALID;
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    QPIC_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    return QPIC_OK;
}

void * qpic_alloc_pinned(size_t bytes) {
    void * p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}

void qpic_free_pinned(void * p) {
    if (p) cudaFreeHost(p);
}

// ---- per-kernel timing of eager slices: "name count total_ms\n" lines, one per kernel, in launch order of first use
int qpic_profile_slices(qpic_ctx * c, int nslices, char * report, size_t report_len) {
    if (!c || !report || report_len == 0) return fail(c, QPIC_ERR_INVALID, "report buffer is NULL");
    if (!c->began) return fail(c, QPIC_ERR_STATE, "qpic_profile_slices called before qpic_begin");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    if (c->lists_dirty) QPIC_TRY(enqueue_rebuild_base_lists(c));
    report[0] = 0;
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    c->profiling = true;
    c->prof_names.clear();
    c->prof_events.clear();
    cudaEvent_t first;
    QPIC_CUDA(c, cudaEventCreate(&first));
    QPIC_CUDA(c, cudaEventRecord(first, c->stream));
    c->prof_names.push_back("(start)");
    c->prof_events.push_back(first);
    int rc = QPIC_OK;
    for (int sidx = 0; sidx < nslices && c->cur > 0 && rc == QPIC_OK; sidx++) {
        rc = enqueue_slice_full(c, c->fast_capable && c->general_left == 0);
        c->cur--;
        c->target = c->cur;
    }
    c->profiling = false;
    if (rc != QPIC_OK) return rc;
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<const char *> names;
    std::vector<double> total;
    std::vector<long> count;
    for (size_t e = 1; e < c->prof_events.size(); e++) {
        float ms = 0.f;
        QPIC_CUDA(c, cudaEventElapsedTime(&ms, c->prof_events[e - 1], c->prof_events[e]));
        size_t k = 0;
        for (; k < names.size(); k++) {
            if (std::strcmp(names[k], c->prof_names[e]) == 0) break;
        }
        if (k == names.size()) {
            names.push_back(c->prof_names[e]);
            total.push_back(0.0);
            count.push_back(0);
        }
        total[k] += ms;
        count[k] += 1;
    }
    for (cudaEvent_t ev : c->prof_events) cudaEventDestroy(ev);
    c->prof_events.clear();
    c->prof_names.clear();
    size_t used = 0;
    for (size_t k = 0; k < names.size(); k++) {
        int w = snprintf(report + used, report_len - used, "%s %ld %.6f\n", names[k], count[k], total[k]);
        if (w < 0 || (size_t)w >= report_len - used) break;
        used += (size_t)w;
    }
    QPIC_TRY(resolve(c));
    return QPIC_OK;
}

// ---- slab-decomposed multi-GPU sweep: mapping the peers' exchange regions
size_t qpic_dist_region_bytes(const qpic_ctx * c) { return (c && c->dist.on) ? c->dist.v.lay.total : 0; }

void * qpic_dist_region(qpic_ctx * c) { return (c && c->dist.on) ? c->dist.region : nullptr; }

int qpic_dist_row_range(const qpic_ctx * c, int * row0, int * nrows) {
    if (!c) return QPIC_ERR_INVALID;
    if (row0) *row0 = row_first(c);
    if (nrows) *nrows = row_count(c);
    return QPIC_OK;
}

int qpic_dist_ipc_handle(qpic_ctx * c, void * handle64) {
#ifdef QPIC_EMU
    (void)handle64;
    return fail(c, QPIC_ERR_UNSUPPORTED, "no IPC in the emulator");
#else
    if (!c || !handle64) return fail(c, QPIC_ERR_INVALID, "handle is NULL");
    if (!c->dist.on) return fail(c, QPIC_ERR_STATE, "not a distributed context");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are 64 bytes");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    cudaIpcMemHandle_t h;
    QPIC_CUDA(c, cudaIpcGetMemHandle(&h, c->dist.region));
    std::memcpy(handle64, &h, sizeof(h));
    return QPIC_OK;
#endif
}

static int dist_finish_attach(qpic_ctx * c) {
    // the view is baked into kernel arguments from here on
    c->dist.attached = true;
    QPIC_CUDA(c, cudaMemsetAsync(c->dist.region + c->dist.v.lay.flags, 0, sizeof(long long) * DIST_MAX, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

int qpic_dist_attach_ipc(qpic_ctx * c, const void * handles) {
#ifdef QPIC_EMU
    (void)handles;
    return fail(c, QPIC_ERR_UNSUPPORTED, "no IPC in the emulator");
#else
    if (!c || !handles) return fail(c, QPIC_ERR_INVALID, "handles is NULL");
    if (!c->dist.on) return fail(c, QPIC_ERR_STATE, "not a distributed context");
    if (c->dist.attached) return fail(c, QPIC_ERR_STATE, "already attached");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    for (int q = 0; q < c->dist.v.world; q++) {
        if (q == c->dist.v.rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char *>(handles) + 64 * (size_t)q, sizeof(h));
        void * mapped = nullptr;
        QPIC_CUDA(c, cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
        c->dist.opened.push_back(mapped);
        c->dist.v.base[q] = static_cast<char *>(mapped);
    }
    return dist_finish_attach(c);
#endif
}

int qpic_dist_attach_pointers(qpic_ctx * c, void * const * regions) {
    if (!c || !regions) return fail(c, QPIC_ERR_INVALID, "regions is NULL");
    if (!c->dist.on) return fail(c, QPIC_ERR_STATE, "not a distributed context");
    if (c->dist.attached) return fail(c, QPIC_ERR_STATE, "already attached");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    for (int q = 0; q < c->dist.v.world; q++) {
        if (q == c->dist.v.rank) continue;
        if (!regions[q]) return fail(c, QPIC_ERR_INVALID, "region of rank %d is NULL", q);
#ifndef QPIC_EMU
        // peers on other devices of this process: enable direct access once (ignored when already on)
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, regions[q]) == cudaSuccess && attr.device != c->prm.device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(attr.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(c, QPIC_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
#endif
        c->dist.v.base[q] = static_cast<char *>(regions[q]);
    }
    return dist_finish_attach(c);
}

// ---- in-graph timeline (only in the -DQPIC_TIMELINE development build, tools/timeline.py): globaltimer stamps that
// block 0 / the last block of every kernel leave behind.  records = int64[4 * cap]: kid, kind, t0, t1 (ns).
// ------------------------------------------------------------------------------------------------ beam particles

int qpic_bunch_set_species(qpic_ctx * c, int count, const double * charge, const double * charge_to_mass_ratio, const int * is_photon) {
    if (!c) return QPIC_ERR_INVALID;
    if (count < 1 || count > BUNCH_MAX_SPECIES || !charge || !charge_to_mass_ratio || !is_photon) {
        return fail(c, QPIC_ERR_INVALID, "species table: 1..%d entries with charge, charge_to_mass_ratio and is_photon", BUNCH_MAX_SPECIES);
    }
    for (int s = 0; s < count; s++) {
        c->bunch.species.charge[s] = charge[s];
        c->bunch.species.charge_to_mass[s] = charge_to_mass_ratio[s];
        c->bunch.species.is_photon[s] = is_photon[s] ? 1 : 0;
    }
    c->bunch.nspecies = count;
    return QPIC_OK;
}

static double * bunch_array(qpic_ctx * c, qpic_bunch_field f) {
    Bunch & b = c->bunch.b;
    double * const table[QPIC_B_COUNT] = {b.x, b.y, b.z, b.px, b.py, b.pz, b.gamma, b.n, b.ex, b.ey, b.ez, b.by, b.bz, b.chi};
    return (f >= 0 && f < QPIC_B_COUNT) ? table[f] : nullptr;
}

int qpic_bunch_init(qpic_ctx * c, size_t count, const double * x, const double * y, const double * z, const double * px, const double * py,
                    const double * pz, const double * gamma, const double * n, const int * species_id) {
    if (!c) return QPIC_ERR_INVALID;
    if (c->dist.on) return fail(c, QPIC_ERR_UNSUPPORTED, "beam particles are not supported in a slab-decomposed sweep");
    if (c->bunch.nspecies == 0) return fail(c, QPIC_ERR_STATE, "qpic_bunch_set_species has to be called before qpic_bunch_init");
    if (count > 0 && (!x || !y || !z || !px || !py || !pz || !gamma || !n || !species_id)) return fail(c, QPIC_ERR_INVALID, "beam particle arrays are NULL");
    if (count > (size_t)std::numeric_limits<int>::max() - 1024) return fail(c, QPIC_ERR_INVALID, "too many beam particles for 32-bit indices");
    const size_t n3 = (size_t)c->g.nx * c->g.n2;
    if (n3 > (size_t)std::numeric_limits<int>::max() - SCAN_CHUNK) return fail(c, QPIC_ERR_UNSUPPORTED, "grid too large for the 32-bit cell keys of the beam deposit");
    for (size_t p = 0; p < count; p++) {
        if (species_id[p] < 0 || species_id[p] >= c->bunch.nspecies) return fail(c, QPIC_ERR_INVALID, "beam particle %zu has species id %d outside the species table", p, species_id[p]);
    }
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    // rho_bunch / jx_bunch enter the psi sources of the slice graph: their addresses have to exist when it is captured
    if ((c->graph_exec || c->graph_exec_fast) && (!c->f3[QPIC_F3_RHO_BUNCH] || !c->f3[QPIC_F3_JX_BUNCH])) {
        return fail(c, QPIC_ERR_STATE, "beam particles must be set up before the first qpic_step of a context");
    }
    if (!c->f3[QPIC_F3_RHO_BUNCH]) QPIC_TRY(dev_alloc(c, &c->f3[QPIC_F3_RHO_BUNCH], n3));
    if (!c->f3[QPIC_F3_JX_BUNCH]) QPIC_TRY(dev_alloc(c, &c->f3[QPIC_F3_JX_BUNCH], n3));
    qpic_ctx::BunchState & S = c->bunch;
    if ((int)count > S.capacity || !S.b.x) {
        const size_t cap = count ? count : 1;
        double ** arrays[QPIC_B_COUNT] = {&S.b.x, &S.b.y, &S.b.z, &S.b.px, &S.b.py, &S.b.pz, &S.b.gamma, &S.b.n, &S.b.ex, &S.b.ey, &S.b.ez, &S.b.by, &S.b.bz, &S.b.chi};
        for (double ** a : arrays) QPIC_TRY(dev_alloc(c, a, cap));
        QPIC_TRY(dev_alloc(c, &S.b.species, cap));
        QPIC_TRY(dev_alloc(c, &S.cell, cap));
        QPIC_TRY(dev_alloc(c, &S.tmp, cap));
        QPIC_TRY(dev_alloc(c, &S.perm, cap));
        S.capacity = (int)cap;
    }
    if (!S.count) {
        S.nchunks = (int)((n3 + SCAN_CHUNK - 1) / SCAN_CHUNK);
        QPIC_TRY(dev_alloc(c, &S.count, n3));
        QPIC_TRY(dev_alloc(c, &S.start, n3 + 1));
        QPIC_TRY(dev_alloc(c, &S.sums, (size_t)S.nchunks));
    }
    S.b.np = (int)count;
    if (count) {
        const double * src[8] = {x, y, z, px, py, pz, gamma, n};
        double * dst[8] = {S.b.x, S.b.y, S.b.z, S.b.px, S.b.py, S.b.pz, S.b.gamma, S.b.n};
        for (int a = 0; a < 8; a++) QPIC_CUDA(c, cudaMemcpyAsync(dst[a], src[a], sizeof(double) * count, cudaMemcpyHostToDevice, c->stream));
        QPIC_CUDA(c, cudaMemcpyAsync(S.b.species, species_id, sizeof(int) * count, cudaMemcpyHostToDevice, c->stream));
        double * zeroed[6] = {S.b.ex, S.b.ey, S.b.ez, S.b.by, S.b.bz, S.b.chi};  // bunch_particle_3d's defaults (src/containers.h:54-59)
        for (double * a : zeroed) QPIC_CUDA(c, cudaMemsetAsync(a, 0, sizeof(double) * count, c->stream));
    }
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

size_t qpic_bunch_count(const qpic_ctx * c) { return c ? (size_t)c->bunch.b.np : 0; }

int qpic_bunch_deposit_densities(qpic_ctx * c) {
    if (!c) return QPIC_ERR_INVALID;
    qpic_ctx::BunchState & S = c->bunch;
    if (!S.count) return fail(c, QPIC_ERR_STATE, "qpic_bunch_deposit_densities called before qpic_bunch_init");
    if (c->began && !c->finished) return fail(c, QPIC_ERR_STATE, "the bunch densities cannot change during a sweep");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    cudaStream_t st = c->stream;
    const int n3 = c->g.nx * c->g.n2;
    const int np = S.b.np;
    QPIC_CUDA(c, cudaMemsetAsync(S.count, 0, sizeof(int) * (size_t)n3, st));
    if (np > 0) {
        QPIC_LAUNCH_FLAT(k_bunch_locate_count, blocks_for(np, 256), 256, 0, st, S.b, c->g, S.cell, S.count);
        QPIC_TRY(check_launch(c, "k_bunch_locate_count"));
    }
    QPIC_LAUNCH(k_scan_chunk_sums, S.nchunks, SCAN_THREADS, 0, st, c->ctl_scratch, S.count, n3, S.sums);
    QPIC_TRY(check_launch(c, "k_scan_chunk_sums"));
    QPIC_LAUNCH(k_scan_sums, 1, SCAN_THREADS, 0, st, c->ctl_scratch, S.sums, S.nchunks);
    QPIC_TRY(check_launch(c, "k_scan_sums"));
    QPIC_LAUNCH(k_bunch_scan_apply, S.nchunks, SCAN_THREADS, 0, st, S.count, n3, S.sums, S.start);
    QPIC_TRY(check_launch(c, "k_bunch_scan_apply"));
    if (np > 0) {
        QPIC_LAUNCH_FLAT(k_bunch_fill, blocks_for(np, 256), 256, 0, st, np, S.cell, S.count, S.tmp);
        QPIC_TRY(check_launch(c, "k_bunch_fill"));
        QPIC_LAUNCH_FLAT(k_bunch_rank, blocks_for(np, 256), 256, 0, st, np, S.cell, S.start, S.tmp, S.perm);
        QPIC_TRY(check_launch(c, "k_bunch_rank"));
    }
    int blocks = blocks_for((size_t)n3, 256);
    if (blocks > 148 * 64) blocks = 148 * 64;
    QPIC_LAUNCH_FLAT(k_bunch_deposit, blocks, 256, 0, st, c->g, S.b, S.species, S.start, S.perm, c->f3[QPIC_F3_RHO_BUNCH], c->f3[QPIC_F3_JX_BUNCH]);
    QPIC_TRY(check_launch(c, "k_bunch_deposit"));
    QPIC_CUDA(c, cudaStreamSynchronize(st));
    return QPIC_OK;
}

int qpic_bunch_gather_fields(qpic_ctx * c, double field_schwinger) {
    if (!c) return QPIC_ERR_INVALID;
    qpic_ctx::BunchState & S = c->bunch;
    if (!S.count) return fail(c, QPIC_ERR_STATE, "qpic_bunch_gather_fields called before qpic_bunch_init");
    if (c->began && !c->finished) return fail(c, QPIC_ERR_STATE, "qpic_bunch_gather_fields called during a sweep (qpic_finish first)");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    if (S.b.np > 0) {
        QPIC_LAUNCH_FLAT(k_bunch_gather, blocks_for(S.b.np, 128), 128, 0, c->stream, S.b, c->g, c->f3[QPIC_F3_EX], c->f3[QPIC_F3_EY], c->f3[QPIC_F3_EZ],
                         c->f3[QPIC_F3_BY], c->f3[QPIC_F3_BZ], field_schwinger);
        QPIC_TRY(check_launch(c, "k_bunch_gather"));
    }
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

int qpic_bunch_advance(qpic_ctx * c, double dt) {
    if (!c) return QPIC_ERR_INVALID;
    qpic_ctx::BunchState & S = c->bunch;
    if (!S.count) return fail(c, QPIC_ERR_STATE, "qpic_bunch_advance called before qpic_bunch_init");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    if (S.b.np > 0) {
        QPIC_LAUNCH_FLAT(k_bunch_advance, blocks_for(S.b.np, 256), 256, 0, c->stream, S.b, S.species, dt);
        QPIC_TRY(check_launch(c, "k_bunch_advance"));
    }
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

int qpic_bunch_download(qpic_ctx * c, qpic_bunch_field f, double * dst) {
    if (!c) return QPIC_ERR_INVALID;
    double * src = bunch_array(c, f);
    if (!src) return fail(c, QPIC_ERR_INVALID, "bad beam particle field id, or no beam particles were set up");
    if (c->bunch.b.np == 0) return QPIC_OK;
    if (!dst) return fail(c, QPIC_ERR_INVALID, "dst is NULL");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    QPIC_CUDA(c, cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)c->bunch.b.np, cudaMemcpyDeviceToHost, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

int qpic_bunch_upload(qpic_ctx * c, qpic_bunch_field f, const double * src) {
    if (!c) return QPIC_ERR_INVALID;
    double * dst = bunch_array(c, f);
    if (!dst || !src) return fail(c, QPIC_ERR_INVALID, "bad beam particle field id / NULL src, or no beam particles were set up");
    if (c->bunch.b.np == 0) return QPIC_OK;
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    QPIC_CUDA(c, cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)c->bunch.b.np, cudaMemcpyHostToDevice, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

// advance_bunch_particles with qed == true (src/System_3d.cpp:862-921)
int qpic_bunch_advance_qed(qpic_ctx * c, double dt, double frequency, const double * u_emit, const double * u_accept, unsigned long long seed,
                           double * max_radiation_probability, size_t * photon_count) {
    if (!c) return QPIC_ERR_INVALID;
    qpic_ctx::BunchState & S = c->bunch;
    if (!S.count) return fail(c, QPIC_ERR_STATE, "qpic_bunch_advance_qed called before qpic_bunch_init");
    if ((u_emit == nullptr) != (u_accept == nullptr)) return fail(c, QPIC_ERR_INVALID, "u_emit and u_accept must be given together (or both NULL: device RNG)");
    if (!(frequency > 0.0)) return fail(c, QPIC_ERR_INVALID, "the base plasma frequency must be positive");
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    cudaStream_t st = c->stream;
    const int np = S.b.np;
    if (S.qed_capacity < S.capacity) {
        const size_t cap = (size_t)S.capacity;
        double ** arrays[] = {&S.photons.x, &S.photons.y, &S.photons.z, &S.photons.px, &S.photons.py, &S.photons.pz, &S.photons.gamma, &S.photons.n,
                              &S.qed.qx, &S.qed.qy, &S.qed.qz, &S.u_emit, &S.u_accept};
        for (double ** a : arrays) QPIC_TRY(dev_alloc(c, a, cap));
        QPIC_TRY(dev_alloc(c, &S.qed.flag, cap));
        QPIC_TRY(dev_alloc(c, &S.qed_start, cap + 1));
        QPIC_TRY(dev_alloc(c, &S.qed_sums, (cap + SCAN_CHUNK - 1) / SCAN_CHUNK));
        QPIC_TRY(dev_alloc(c, &S.qed.max_bits, 1));
        S.qed_capacity = S.capacity;
    }
    S.photons.np = 0;
    double maxp = 0.0;
    if (np > 0) {
        if (u_emit) {
            QPIC_CUDA(c, cudaMemcpyAsync(S.u_emit, u_emit, sizeof(double) * np, cudaMemcpyHostToDevice, st));
            QPIC_CUDA(c, cudaMemcpyAsync(S.u_accept, u_accept, sizeof(double) * np, cudaMemcpyHostToDevice, st));
        }
        QPIC_CUDA(c, cudaMemsetAsync(S.qed.max_bits, 0, sizeof(unsigned long long), st));
        // W_new's constants with the reference's expressions (src/monte_carlo.cpp:17, src/constants.h:3-12)
        const double PI = 3.141592653589793238, ELECTRON_MASS_CGS = 9.1093837015e-28, PLANCK_CONST_CGS = 6.62607015e-27;
        const double PLANCK_CONST_BAR_CGS = PLANCK_CONST_CGS / (2 * PI), SPEED_OF_LIGHT_CGS = 29979245800.0, ELEMENTARY_CHARGE_CGS = 4.803204672997e-10;
        QedConstants Q;
        Q.coef = ELECTRON_MASS_CGS * SPEED_OF_LIGHT_CGS * SPEED_OF_LIGHT_CGS / (PLANCK_CONST_BAR_CGS * frequency);
        Q.fine_structure = (ELEMENTARY_CHARGE_CGS * ELEMENTARY_CHARGE_CGS / PLANCK_CONST_BAR_CGS / SPEED_OF_LIGHT_CGS);
        QPIC_LAUNCH_FLAT(k_bunch_advance_qed, blocks_for(np, QPIC_PBLOCK), QPIC_PBLOCK, 0, st, S.b, S.species, dt, Q, u_emit ? (const double *)S.u_emit : (const double *)nullptr,
                         u_emit ? (const double *)S.u_accept : (const double *)nullptr, seed, S.qed_step, S.qed);
        QPIC_TRY(check_launch(c, "k_bunch_advance_qed"));
        const int nchunks = (np + SCAN_CHUNK - 1) / SCAN_CHUNK;
        QPIC_LAUNCH(k_scan_chunk_sums, nchunks, SCAN_THREADS, 0, st, c->ctl_scratch, S.qed.flag, np, S.qed_sums);
        QPIC_TRY(check_launch(c, "k_scan_chunk_sums"));
        QPIC_LAUNCH(k_scan_sums, 1, SCAN_THREADS, 0, st, c->ctl_scratch, S.qed_sums, nchunks);
        QPIC_TRY(check_launch(c, "k_scan_sums"));
        QPIC_LAUNCH(k_bunch_scan_apply, nchunks, SCAN_THREADS, 0, st, S.qed.flag, np, S.qed_sums, S.qed_start);
        QPIC_TRY(check_launch(c, "k_bunch_scan_apply"));
        Bunch ph = S.photons;
        ph.np = np;
        QPIC_LAUNCH_FLAT(k_bunch_photon_compact, blocks_for(np, 256), 256, 0, st, S.b, S.qed, S.qed_start, ph);
        QPIC_TRY(check_launch(c, "k_bunch_photon_compact"));
        int emitted = 0;
        unsigned long long bits = 0;
        QPIC_CUDA(c, cudaMemcpyAsync(&emitted, S.qed_start + np, sizeof(int), cudaMemcpyDeviceToHost, st));
        QPIC_CUDA(c, cudaMemcpyAsync(&bits, S.qed.max_bits, sizeof(bits), cudaMemcpyDeviceToHost, st));
        QPIC_CUDA(c, cudaStreamSynchronize(st));
        S.photons.np = emitted;
        std::memcpy(&maxp, &bits, sizeof(maxp));
    }
    S.qed_step++;
    if (max_radiation_probability) *max_radiation_probability = maxp;
    if (photon_count) *photon_count = (size_t)S.photons.np;
    return QPIC_OK;
}

size_t qpic_photon_count(const qpic_ctx * c) { return c ? (size_t)c->bunch.photons.np : 0; }

int qpic_photons_download(qpic_ctx * c, qpic_bunch_field f, double * dst) {
    if (!c || f < 0 || f >= QPIC_B_COUNT) return fail(c, QPIC_ERR_INVALID, "bad photon field id");
    const Bunch & ph = c->bunch.photons;
    const size_t n = (size_t)ph.np;
    if (n == 0) return QPIC_OK;
    if (!dst) return fail(c, QPIC_ERR_INVALID, "dst is NULL");
    const double * const table[QPIC_B_COUNT] = {ph.x, ph.y, ph.z, ph.px, ph.py, ph.pz, ph.gamma, ph.n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (!table[f]) {  // ex, ey, ez, by, bz, chi of a freshly emitted photon: bunch_particle_3d's defaults (src/containers.h:54-59)
        std::memset(dst, 0, sizeof(double) * n);
        return QPIC_OK;
    }
    QPIC_CUDA(c, cudaSetDevice(c->prm.device));
    QPIC_CUDA(c, cudaMemcpyAsync(dst, table[f], sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    QPIC_CUDA(c, cudaStreamSynchronize(c->stream));
    return QPIC_OK;
}

int qpic_timeline_reset(qpic_ctx * c) {
#ifdef QPIC_TIMELINE
    const int zero = 0;
    QPIC_CUDA(c, cudaDeviceSynchronize());
    QPIC_CUDA(c, cudaMemcpyToSymbol(g_tl_n, &zero, sizeof(int)));
    return QPIC_OK;
#else
    return fail(c, QPIC_ERR_STATE, "not a timeline build (compile with -DQPIC_TIMELINE)");
#endif
}

long qpic_timeline_read(qpic_ctx * c, long long * records, long cap) {
#ifdef QPIC_TIMELINE
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    int n = 0;
    if (cudaMemcpyFromSymbol(&n, g_tl_n, sizeof(int)) != cudaSuccess) return -1;
    if (n > TL_CAPACITY) n = TL_CAPACITY;
    if (n > cap) n = (int)cap;
    std::vector<TlRecord> host((size_t)n);
    if (n > 0 && cudaMemcpyFromSymbol(host.data(), g_tl, sizeof(TlRecord) * (size_t)n) != cudaSuccess) return -1;
    for (int i = 0; i < n; i++) {
        records[4 * i + 0] = host[i].kid;
        records[4 * i + 1] = host[i].kind;
        records[4 * i + 2] = (long long)host[i].t0;
        records[4 * i + 3] = (long long)host[i].t1;
    }
    return n;
#else
    (void)records;
    (void)cap;
    fail(c, QPIC_ERR_STATE, "not a timeline build (compile with -DQPIC_TIMELINE)");
    return -1;
#endif
}

// name of a kernel id of the timeline records ("" when out of range)
const char * qpic_timeline_kernel_name(int kid) {
    static const char * names[] = QPIC_KERNEL_NAMES;
    static const char * tile[] = {"k_tile_deposit_charge", "k_tile_deposit_charge_mid", "k_tile_deposit_jx", "?", "?", "k_tile_deposit_jy_jz", "?", "?", "?", "?",
                                  "k_tile_deposit_rho_chi", "k_tile_fold_rho_chi", "?", "?", "?", "?", "k_p2_cols_solve_complex"};
    if (kid >= 0 && kid < KID_COUNT) return names[kid];
    if (kid >= KID_COUNT && kid <= KID_COUNT + 16) return tile[kid - KID_COUNT];
    return "";
}

}  // extern "C"
