// Particle / grid kernels of the per-xi-slice solver (sm_100a, FP64).  Everything here is HBM/L2-bound FP64 work:
// no tensor cores.  Citations are to /root/reference/src (file:line) -- the behaviour each kernel reproduces.
//
// Data layout
//   2-D fields: double[ny*nz], z fastest (array2d, src/array2d.h:15-16).  3-D fields: double[nx*ny*nz], x slowest.
//   Plasma macro-particles: SoA of doubles (the reference's AoS wake_particle_2d, src/containers.h:31-43), plus a
//   per-particle "cell descriptor" (cell = jc*nz+kc, fy, fz) = the (int)(y/dy) / fractional split every deposit and
//   gather of the reference starts with (src/array_utils.cpp:121-127).  Staggered indices are derived from it.
//
// Deterministic deposition: particles are counting-sorted by cell (index lists, canonical order inside a cell),
// then ONE THREAD PER GRID NODE sums the contributions of the particles in the neighbouring cells in a fixed order
// -- no floating-point atomics anywhere, so every run gives identical bits.
#pragma once

#include "qpic_compat.h"

namespace qpic {

// ------------------------------------------------------------------------------------------------ kernel prologue
// Every kernel starts with QPIC_KERNEL_BEGIN(id):
//   * griddepcontrol.wait -- the kernels of a slice are launched with programmatic stream serialization (PDL), so a
//     kernel's blocks may become resident while its predecessor in the stream is still running; nothing produced by an
//     earlier kernel (including the control block) may be read before this point.  Without the launch attribute the
//     instruction is a no-op.
//   * griddepcontrol.launch_dependents right after it: the successor may be staged as soon as all blocks of this kernel
//     have passed their own wait (i.e. at most one waiting kernel is resident at a time).
//   * -DQPIC_TIMELINE (tools/timeline.py, separate library): block 0 records globaltimer stamps before / after the wait and
//     the last block to leave records the end, giving the in-graph timeline of a slice (gaps, overlaps) that ncu cannot.
enum KernelId {
    KID_k_init_a_sqr,
    KID_k_locate_count,
    KID_k_scan_chunk_sums,
    KID_k_scan_sums,
    KID_k_scan_apply,
    KID_k_sort_fill,
    KID_k_sort_rank,
    KID_k_deposit_charge,
    KID_k_deposit_jx_rho_chi,
    KID_k_deposit_jy_jz,
    KID_k_potential_fixup,
    KID_k_particles_to_slots,
    KID_k_particles_gamma,
    KID_k_particles_predict,
    KID_k_particles_currents_yz,
    KID_k_particles_advance,
    KID_k_end_of_slice,
    KID_k_next_slice,
    KID_k_tile_deposit,
    KID_k_fft_rows_r2c,
    KID_k_fft_rows_c2r,
    KID_k_fft_cols_solve,
    KID_k_fft_rows_b_forward,
    KID_k_fft_rows_b_backward,
    KID_k_fft_rows_c2c,
    KID_k_p2_rows_r2c,
    KID_k_p2_rows_c2r,
    KID_k_p2_cols_solve,
    KID_k_p2_rows_b_forward,
    KID_k_p2_rows_b_backward,
    KID_k_p2_rows_c2c,
    KID_k_xchg_barrier,
    KID_k_halo_reduce_send,
    KID_k_halo_reduce_add,
    KID_k_halo_fill,
    KID_k_migrate_publish,
    KID_k_migrate_unpack,
    KID_k_xy_cut,
    KID_k_xy_lines,
    KID_k_bunch_locate_count,
    KID_k_bunch_scan_apply,
    KID_k_bunch_fill,
    KID_k_bunch_rank,
    KID_k_bunch_deposit,
    KID_k_bunch_gather,
    KID_k_bunch_advance,
    KID_k_bunch_advance_qed,
    KID_k_bunch_photon_compact,
    KID_COUNT
};
#define QPIC_KERNEL_NAMES { "k_init_a_sqr", "k_locate_count", "k_scan_chunk_sums", "k_scan_sums", "k_scan_apply", "k_sort_fill", "k_sort_rank", "k_deposit_charge", "k_deposit_jx_rho_chi", "k_deposit_jy_jz", "k_potential_fixup", "k_particles_to_slots", "k_particles_gamma", "k_particles_predict", "k_particles_currents_yz", "k_particles_advance", "k_end_of_slice", "k_next_slice", "k_tile_deposit", "k_fft_rows_r2c", "k_fft_rows_c2r", "k_fft_cols_solve", "k_fft_rows_b_forward", "k_fft_rows_b_backward", "k_fft_rows_c2c", "k_p2_rows_r2c", "k_p2_rows_c2r", "k_p2_cols_solve", "k_p2_rows_b_forward", "k_p2_rows_b_backward", "k_p2_rows_c2c", "k_xchg_barrier", "k_halo_reduce_send", "k_halo_reduce_add", "k_halo_fill", "k_migrate_publish", "k_migrate_unpack", "k_xy_cut", "k_xy_lines", "k_bunch_locate_count", "k_bunch_scan_apply", "k_bunch_fill", "k_bunch_rank", "k_bunch_deposit", "k_bunch_gather", "k_bunch_advance", "k_bunch_advance_qed", "k_bunch_photon_compact" }

#ifdef QPIC_EMU
#define QPIC_KERNEL_BEGIN(id) ((void)0)
#define QPIC_KERNEL_TAIL() ((void)0)
#else
#define QPIC_KERNEL_TAIL() qpic::pdl_trigger()
__device__ __forceinline__ void pdl_grid_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// QPIC_KERNEL_TAIL(): placed where a kernel's remaining work is short (before its final stores).  Once every block of the grid
// has passed it (or exited) the next kernel of the stream may be staged; that kernel still waits for this one to finish
// completely (griddepcontrol.wait), so the placement is purely a latency knob.  Without it the trigger is the block's exit.
__device__ __forceinline__ void pdl_trigger() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

#ifdef QPIC_TIMELINE
constexpr int TL_CAPACITY = 1 << 16;
struct TlRecord {
    int kid, kind;  // kind 0: start (t0 = block 0 resident, t1 = dependencies satisfied), 1: end (t0)
    unsigned long long t0, t1;
};
__device__ TlRecord g_tl[TL_CAPACITY];
__device__ int g_tl_n;
__device__ unsigned int g_tl_exit[KID_COUNT + 32];
__device__ __forceinline__ unsigned long long tl_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
struct TlSpan {
    int kid;
    __device__ __forceinline__ TlSpan(int kid_) : kid(kid_) {
        const bool first = (threadIdx.x | threadIdx.y | blockIdx.x | blockIdx.y | blockIdx.z) == 0;
        const unsigned long long t0 = first ? tl_now() : 0ull;
        pdl_grid_sync();
        if (first) {
            const int slot = atomicAdd(&g_tl_n, 1);
            if (slot < TL_CAPACITY) g_tl[slot] = TlRecord{kid, 0, t0, tl_now()};
        }
    }
    __device__ __forceinline__ ~TlSpan() {
        if ((threadIdx.x | threadIdx.y) == 0) {
            const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
            const unsigned int n = atomicAdd(&g_tl_exit[kid], 1u) + 1u;
            if (n % total == 0u) {
                const int slot = atomicAdd(&g_tl_n, 1);
                if (slot < TL_CAPACITY) g_tl[slot] = TlRecord{kid, 1, tl_now(), 0ull};
            }
        }
    }
};
#define QPIC_KERNEL_BEGIN(id) qpic::TlSpan qpic_tl_span_(id)
#else
#define QPIC_KERNEL_BEGIN(id) qpic::pdl_grid_sync()
#endif
#endif

// L2 residency hints for what one kernel writes and the next tile kernel reads (the predictor's half-step outputs -> jy/jz deposit): stores marked
// evict-last keep their lines in the 126 MB L2 while the predictor streams ~100 MB of read-once slot arrays past them; the consumer's loads mark
// them evict-first again.  A/B knob QPIC_L2_KEEP (bits): 1 the predictor's half-step outputs are stored evict-last, 2 the jy/jz deposit's loads of them are evict-first,
// 4 the predictor's load of the weights (read again by the jy/jz deposit) is evict-last, 8 its corner-sum stores are.  (sm_100: the plain .L2::evict_* qualifiers need
// 256-bit accesses, so the hint travels as a createpolicy descriptor.)
#ifndef QPIC_L2_KEEP
#define QPIC_L2_KEEP 1
#endif
#if QPIC_L2_KEEP && !defined(QPIC_EMU)
__device__ __forceinline__ unsigned long long l2_policy_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_keep(double2 * p, double2 v) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(l2_policy_last()) : "memory");
}
__device__ __forceinline__ void st_keep1(double * p, double v) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(l2_policy_last()) : "memory");
}
__device__ __forceinline__ double ld_keep1(const double * p) {
    double r;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(l2_policy_last()));
    return r;
}
__device__ __forceinline__ void st_keep(int * p, int v) {
    asm volatile("st.global.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(l2_policy_last()) : "memory");
}
#else
__device__ __forceinline__ void st_keep(double2 * p, double2 v) { *p = v; }
__device__ __forceinline__ void st_keep(int * p, int v) { *p = v; }
__device__ __forceinline__ void st_keep1(double * p, double v) { *p = v; }
__device__ __forceinline__ double ld_keep1(const double * p) { return *p; }
#endif
#if (QPIC_L2_KEEP & 2) && !defined(QPIC_EMU)
__device__ __forceinline__ double2 ld_release(const double2 * p) {
    double2 r;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(l2_policy_first()));
    return r;
}
__device__ __forceinline__ int ld_release(const int * p) {
    int r;
    asm volatile("ld.global.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(l2_policy_first()));
    return r;
}
#else
__device__ __forceinline__ double2 ld_release(const double2 * p) { return *p; }
__device__ __forceinline__ int ld_release(const int * p) { return *p; }
#endif

// streaming load: the line is marked evict-first in L2.  For slot arrays that are read once per iteration, so that they do not
// push out the arrays the deposit kernels are about to read.
__device__ __forceinline__ double ld_stream(const double * p) {
#if defined(QPIC_NO_STREAM_HINT) || defined(QPIC_EMU)
    return *p;
#else
    return __ldcs(p);
#endif
}
__device__ __forceinline__ double2 ld_stream2(const double2 * p) {
#if defined(QPIC_NO_STREAM_HINT) || defined(QPIC_EMU)
    return *p;
#else
    return __ldcs(p);
#endif
}
// the opposite hint: evict-last, for slot arrays that every B iteration of the slice reads again (A/B knob QPIC_PIN_MASK: bit i
// = i-th slice invariant of k_particles_predict keeps its L2 lines; 0 = all streamed).  Measured on B200, config #2
// (gpurun_out/r31_pin.log): mask 0 0.767 ms/slice, 0x3 0.771, 0x1f 0.786, 0x3ff 0.797 -- pinned invariants push the predictor's
// outputs out of L2 and the latency-bound deposits that read them next lose more than the predictor gains.  Stays 0.
#ifndef QPIC_PIN_MASK
#define QPIC_PIN_MASK 0
#endif
template <int BIT>
__device__ __forceinline__ double ld_invariant(const double * p) {
#if defined(QPIC_EMU)
    return *p;
#else
    if ((QPIC_PIN_MASK >> BIT) & 1) {
        double v;
        unsigned long long policy;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
        asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy));
        return v;
    }
    return ld_stream(p);
#endif
}
struct Ctl {
    int slice;      // xi-slice being solved (nx-1 ... 0)
    int stop_tag;   // 0, or slice * STOP_TAG_STRIDE + iteration + 1 of the B-field solve that tripped the stop condition
    int fast_fail;  // != 0: a half-step displacement left the +-1 cell window of the sort-free deposit; the fast-path
                    // kernels of this and later slices are no-ops until the host redoes the slice on the general path
    int dist_error; // slab-decomposed runs: != 0 after a bounded exchange spin gave up or an exchange buffer overflowed
    // Slab-decomposed runs migrate particles between GPUs, so the particle counts live on the device (count + 1; 0 = use
    // the static count the kernel was launched with -- what a zero-initialised control block of a single-GPU run gives):
    int np1;        // live slots of this slice
    int np_p1;      // entries of the particle-id arrays (live + emigrated + just arrived) the cell sort has to look at
    int mig_base1;  // first entry of this slice's arrivals + 1 (k_migrate_unpack; 0: none), see sort_key
    int pad;
};

__device__ __forceinline__ int live_count(const Ctl * ctl, int np_static) {
    const int v = ctl->np1;
    return v ? v - 1 : np_static;
}
__device__ __forceinline__ int entry_count(const Ctl * ctl, int np_static) {
    const int v = ctl->np_p1;
    return v ? v - 1 : np_static;
}

// every kernel of the xi-loop starts with this test (the B backward rows have their own, see qpic_fft.cuh)
__device__ __forceinline__ bool cancelled(const Ctl * ctl) { return (ctl->stop_tag | ctl->fast_fail) != 0; }
constexpr int STOP_TAG_STRIDE = 1024;

struct Geom {
    int nx, ny, nz;
    int n2;  // ny * nz
    double dx, dy, dz;
    double ly, lz;
    // 1/dx, 1/dy, 1/dz: finite differences multiply by these instead of dividing (exact when the step is a power of two,
    // as in the BASELINE grids; otherwise within an ulp of the reference's quotient).  Positions -> cells keep the division.
    double inv_dx, inv_dy, inv_dz;
    // dy / dz is a power of two (all BASELINE grids): y * inv_dy IS y / dy bit for bit, so locate_fast may multiply (an FP64 division
    // costs ~30 instructions; the particle kernels spent most of their issue slots on divisions)
    int exact_dy, exact_dz;
};

// ------------------------------------------------------------------------------------------------ index helpers

// (int)(y/d) and the fractional part, as every reference deposit / gather begins (src/array_utils.cpp:121-127; for
// y >= 0 identical to floor((y - 0)/d), :90-96).  y within an ulp of l can round y/d up to n: the reference would
// index out of bounds there (asserts are compiled out); we map it to the same periodic point, cell 0.
__device__ __forceinline__ void locate(double y, double d, int n, int & jc, double & f) {
    const double t = y / d;
    jc = (int)t;
    f = t - (double)jc;
    if (jc >= n) {
        jc = 0;
        f = 0.0;
    }
}

// the same with the quotient formed as y * (1/d) where that is exact (d a power of two); otherwise the division
__device__ __forceinline__ void locate_fast(double y, double d, double inv_d, int exact, int n, int & jc, double & f) {
    const double t = exact ? y * inv_d : y / d;
    jc = (int)t;
    f = t - (double)jc;
    if (jc >= n) {
        jc = 0;
        f = 0.0;
    }
}

// index / fraction on a grid whose origin is shifted by +d/2 (jy, jz, by, bz staggering, src/System_3d.cpp:180-193;
// derivative gathers' floor(y/d - 0.5), src/array_utils.cpp:194-196)
__device__ __forceinline__ void stagger(int jc, double f, int n, int & j1, double & fs) {
    if (f >= 0.5) {
        j1 = jc;
        fs = f - 0.5;
    } else {
        j1 = jc - 1;
        if (j1 < 0) j1 += n;
        fs = f + 0.5;
    }
}

__device__ __forceinline__ int wrap_inc(int j, int n) {  // (j + 1) % n for 0 <= j < n
    j += 1;
    return j >= n ? j - n : j;
}

__device__ __forceinline__ int wrap_dec(int j, int n) {  // (j - 1 + n) % n
    return j == 0 ? n - 1 : j - 1;
}

// 4-point CIC gather given first index / fraction per axis (src/array_utils.cpp:167-188)
__device__ __forceinline__ double gather4(const double * __restrict__ a, int nz, int j1, int j2, int k1, int k2, double fy, double fz) {
    return a[nz * j1 + k1] * (1 - fy) * (1 - fz) + a[nz * j2 + k1] * fy * (1 - fz) + a[nz * j1 + k2] * (1 - fy) * fz +
           a[nz * j2 + k2] * fy * fz;
}

// node-centred field at the particle
__device__ __forceinline__ double gather_node(const double * __restrict__ a, int ny, int nz, int jc, int kc, double fy, double fz) {
    return gather4(a, nz, jc, wrap_inc(jc, ny), kc, wrap_inc(kc, nz), fy, fz);
}

// d/dy at half points: 6-cell stencil (src/array_utils.cpp:240-263 / :190-213)
__device__ __forceinline__ double gather_yder(const double * __restrict__ a, int ny, int nz, double inv_dy, int jc, int kc, double fy, double fz) {
    int j1;
    double fs;
    stagger(jc, fy, ny, j1, fs);
    const int j2 = wrap_inc(j1, ny), j3 = wrap_inc(j2, ny);
    const int k1 = kc, k2 = wrap_inc(kc, nz);
    const double d11 = (a[nz * j2 + k1] - a[nz * j1 + k1]) * inv_dy;
    const double d21 = (a[nz * j3 + k1] - a[nz * j2 + k1]) * inv_dy;
    const double d12 = (a[nz * j2 + k2] - a[nz * j1 + k2]) * inv_dy;
    const double d22 = (a[nz * j3 + k2] - a[nz * j2 + k2]) * inv_dy;
    return d11 * (1 - fs) * (1 - fz) + d21 * fs * (1 - fz) + d12 * (1 - fs) * fz + d22 * fs * fz;
}

// d/dz at half points (src/array_utils.cpp:265-288 / :215-238)
__device__ __forceinline__ double gather_zder(const double * __restrict__ a, int ny, int nz, double inv_dz, int jc, int kc, double fy, double fz) {
    int k1;
    double fs;
    stagger(kc, fz, nz, k1, fs);
    const int k2 = wrap_inc(k1, nz), k3 = wrap_inc(k2, nz);
    const int j1 = jc, j2 = wrap_inc(jc, ny);
    const double d11 = (a[nz * j1 + k2] - a[nz * j1 + k1]) * inv_dz;
    const double d21 = (a[nz * j2 + k2] - a[nz * j2 + k1]) * inv_dz;
    const double d12 = (a[nz * j1 + k3] - a[nz * j1 + k2]) * inv_dz;
    const double d22 = (a[nz * j2 + k3] - a[nz * j2 + k2]) * inv_dz;
    return d11 * (1 - fy) * (1 - fs) + d21 * fy * (1 - fs) + d12 * (1 - fy) * fs + d22 * fy * fs;
}

// src/System_3d.cpp:668-684
__device__ __forceinline__ double wrap_coordinate(double y, double l) {
    y = fmod(y, l);
    if (y < 0) y += l;
    return y;
}

// ------------------------------------------------------------------------------------------------ setup kernels

// a_sqr = amplitude * exp(-(x-x0)^2/wx/wx - (y-y0)^2/wy/wy - (z-z0)^2/wz/wz), node centred
// (src/System_3d.cpp:805-816, src/profiles.cpp:4-11)
__global__ void k_init_a_sqr(Geom g, double * __restrict__ a, double amp, double wx, double wy, double wz, double x0, double y0, double z0) {
    QPIC_KERNEL_BEGIN(KID_k_init_a_sqr);
    const size_t total = (size_t)g.nx * g.n2;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(t % g.nz);
        const int j = (int)((t / g.nz) % g.ny);
        const int i = (int)(t / g.n2);
        const double x = i * g.dx, y = j * g.dy, z = k * g.dz;
        a[t] = amp * exp(-(x - x0) * (x - x0) / wx / wx - (y - y0) * (y - y0) / wy / wy - (z - z0) * (z - z0) / wz / wz);
    }
}

// cell descriptors of (y, z) + per-cell population count (integer atomics: order independent)
__global__ void k_locate_count(int np, Geom g, const double * __restrict__ y, const double * __restrict__ z, int * __restrict__ cell,
                               double * __restrict__ fy, double * __restrict__ fz, int * __restrict__ count) {
    QPIC_KERNEL_BEGIN(KID_k_locate_count);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= np) return;
    int jc, kc;
    double f1, f2;
    locate(y[p], g.dy, g.ny, jc, f1);
    locate(z[p], g.dz, g.nz, kc, f2);
    const int c = jc * g.nz + kc;
    cell[p] = c;
    fy[p] = f1;
    fz[p] = f2;
    atomicAdd(&count[c], 1);
}

// ------------------------------------------------------------------------------------------------ counting sort by cell
// count[] -> start[] (exclusive prefix sum), cursor[] = start[], count[] reset to 0 for the next use.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

// exclusive prefix of v over the 256 threads of the block; total = block sum (valid in all threads)
__device__ __forceinline__ int block_exclusive_scan(int v, int & total) {
    __shared__ int warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // protects warp_sums against the previous call
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    int offset = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        const int s = warp_sums[w];
        if (w < wid) offset += s;
        tot += s;
    }
    total = tot;
    return offset + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_chunk_sums(const Ctl * ctl, const int * __restrict__ count, int n, int * __restrict__ sums) {
    QPIC_KERNEL_BEGIN(KID_k_scan_chunk_sums);
    if (cancelled(ctl)) return;
    const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    int v = 0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; t++) {
        if (base + t < n) v += count[base + t];
    }
    int total;
    block_exclusive_scan(v, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single block: exclusive scan of the chunk sums in place
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(const Ctl * ctl, int * __restrict__ sums, int nchunks) {
    QPIC_KERNEL_BEGIN(KID_k_scan_sums);
    if (cancelled(ctl)) return;
    int carry = 0;
    for (int base = 0; base < nchunks; base += SCAN_THREADS) {
        const int idx = base + threadIdx.x;
        const int v = idx < nchunks ? sums[idx] : 0;
        int total;
        const int ex = block_exclusive_scan(v, total);
        if (idx < nchunks) sums[idx] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
    k_scan_apply(const Ctl * ctl, int * __restrict__ count, int n, const int * __restrict__ sums, int * __restrict__ start, int * __restrict__ cursor,
                 int * __restrict__ live_out, int fold_sums) {
    QPIC_KERNEL_BEGIN(KID_k_scan_apply);
    if (cancelled(ctl)) return;
    // fold_sums: sums[] still holds the raw chunk totals (k_scan_sums was skipped: few chunks) and every block adds up the chunks
    // before its own -- integer sums, so the order is irrelevant
    int chunk_offset;
    if (fold_sums) {
        int part = 0;
        for (int i = threadIdx.x; i < (int)blockIdx.x; i += SCAN_THREADS) part += sums[i];
        block_exclusive_scan(part, chunk_offset);
    } else {
        chunk_offset = sums[blockIdx.x];
    }
    const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
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
            cursor[base + t] = run;
            count[base + t] = 0;
            // slab-decomposed runs: the last bucket collects the emigrated entries, everything before it is live
            if (live_out && base + t == n - 1) *live_out = run + 1;
            run += c[t];
        }
    }
    // the last thread of the last chunk ends on the grand total
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) start[n] = run;
}

// bucket fill: slot order inside a cell is whatever order the integer atomics resolve in ...
__global__ void k_sort_fill(const Ctl * ctl, int np, const int * __restrict__ cell, int * __restrict__ cursor, int * __restrict__ tmp) {
    QPIC_KERNEL_BEGIN(KID_k_sort_fill);
    if (cancelled(ctl)) return;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= entry_count(ctl, np)) return;
    const int slot = atomicAdd(&cursor[cell[p]], 1);
    tmp[slot] = p;
}

// ... so each cell's bucket is put into canonical (ascending particle index) order: rank = number of smaller ids.
// Buckets hold ~ppc entries (a few; tens at strong compression), so the O(k) scan per entry stays in L1.
__global__ void k_sort_rank(const Ctl * ctl, int np, const int * __restrict__ cell, const int * __restrict__ start, const int * __restrict__ tmp,
                            int * __restrict__ perm) {
    QPIC_KERNEL_BEGIN(KID_k_sort_rank);
    if (cancelled(ctl)) return;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= entry_count(ctl, np)) return;
    const int p = tmp[s];
    const int c = cell[p];
    const int b = start[c], e = start[c + 1];
    int rank = 0;
    for (int t = b; t < e; t++) rank += (tmp[t] < p) ? 1 : 0;
    perm[b + rank] = p;
}

// ------------------------------------------------------------------------------------------------ node-centric deposits
// One thread per grid node (j,k).  Node-centred target: the node receives from the particles of cells
// (j,k) [weight (1-fy)(1-fz)], (j-1,k) [fy(1-fz)], (j,k-1) [(1-fy)fz], (j-1,k-1) [fy fz] -- the transpose of the
// 4-point scatter of src/array_utils.cpp:75-82 / :134-141, each product evaluated as (value*wy)*wz like there.

struct CellLists {
    const int * start;  // n2 + 1
    const int * perm;   // np, indices grouped by cell; NULL = identity (the value arrays are already in list order)
    const double2 * fyz;  // (fy, fz) per particle
};

// psi source: fourier.in = jx_bunch(i) - rho_bunch(i) - rho_ion, then deposit of -n (src/System_3d.cpp:342-353, :446-457)
// with_background == 0: plain deposit of -n, i.e. rho_ion itself (src/System_3d.cpp:319-323)
__global__ void k_deposit_charge(const Ctl * ctl, Geom g, CellLists L, const double * __restrict__ nw, const double * __restrict__ rho_ion,
                                 const double * __restrict__ jxb3, const double * __restrict__ rhob3, int with_background,
                                 double * __restrict__ out) {
    QPIC_KERNEL_BEGIN(KID_k_deposit_charge);
    if (cancelled(ctl)) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.n2) return;
    const int j = t / g.nz, k = t - j * g.nz;
    double acc = 0.0;
    if (with_background) {
        const size_t off = (size_t)ctl->slice * g.n2 + t;
        const double jb = jxb3 ? jxb3[off] : 0.0;
        const double rb = rhob3 ? rhob3[off] : 0.0;
        acc = jb - rb - rho_ion[t];
    }
    const int jm = wrap_dec(j, g.ny), km = wrap_dec(k, g.nz);
#pragma unroll
    for (int oy = 0; oy < 2; oy++) {
#pragma unroll
        for (int oz = 0; oz < 2; oz++) {
            const int c = (oy ? jm : j) * g.nz + (oz ? km : k);
            const int e = L.start[c + 1];
            for (int s = L.start[c]; s < e; s++) {
                const int p = L.perm ? L.perm[s] : s;
                const double2 f12 = L.fyz[p];
                const double f1 = f12.x, f2 = f12.y;
                const double wy = oy ? f1 : 1 - f1;
                const double wz = oz ? f2 : 1 - f2;
                acc += -nw[p] * wy * wz;
            }
        }
    }
    out[t] = acc;
}

// jx, rho, susceptibility(i) from per-particle values at (y,z) (src/System_3d.cpp:381-397 and :511-527).  Inside the B
// iterations only jx feeds a solve (:533, :558); rho and susceptibility(i) are overwritten by every later iteration, so
// they are deposited once, from the last executed iteration's values (WITH_RHO_CHI), which is what the reference ends
// up with.  `final_pass`: also runs in the slice where the stop condition tripped (values of the stopping iteration).
template <bool WITH_JX, bool WITH_RHO_CHI>
__global__ void k_deposit_jx_rho_chi(const Ctl * ctl, Geom g, CellLists L, const double * __restrict__ vjx, const double * __restrict__ vrho,
                                     const double * __restrict__ vchi, const double * __restrict__ jxb3, const double * __restrict__ rhob3,
                                     double * __restrict__ jx, double * __restrict__ rho, double * __restrict__ chi3, int final_pass) {
    QPIC_KERNEL_BEGIN(KID_k_deposit_jx_rho_chi);
    if (final_pass) {
        if (ctl->fast_fail || (ctl->stop_tag != 0 && (ctl->stop_tag - 1) / STOP_TAG_STRIDE != ctl->slice)) return;
    } else if (cancelled(ctl)) {
        return;
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.n2) return;
    const int j = t / g.nz, k = t - j * g.nz;
    const size_t off = (size_t)ctl->slice * g.n2 + t;
    double ajx = (WITH_JX && jxb3) ? jxb3[off] : 0.0;
    double arho = (WITH_RHO_CHI && rhob3) ? rhob3[off] : 0.0;
    double achi = 0.0;
    const int jm = wrap_dec(j, g.ny), km = wrap_dec(k, g.nz);
#pragma unroll
    for (int oy = 0; oy < 2; oy++) {
#pragma unroll
        for (int oz = 0; oz < 2; oz++) {
            const int c = (oy ? jm : j) * g.nz + (oz ? km : k);
            const int e = L.start[c + 1];
            for (int s = L.start[c]; s < e; s++) {
                const int p = L.perm ? L.perm[s] : s;
                const double2 f12 = L.fyz[p];
                const double f1 = f12.x, f2 = f12.y;
                const double wy = oy ? f1 : 1 - f1;
                const double wz = oz ? f2 : 1 - f2;
                if (WITH_JX) ajx += vjx[p] * wy * wz;
                if (WITH_RHO_CHI) {
                    arho += vrho[p] * wy * wz;
                    achi += vchi[p] * wy * wz;
                }
            }
        }
    }
    if (WITH_JX) jx[t] = ajx;
    if (WITH_RHO_CHI) {
        rho[t] = arho;
        chi3[off] = achi;
    }
}

// Weight of a particle with unstaggered fraction f, sitting `role` cells (-1, 0, +1) away from the node's own index,
// towards a node of the grid shifted by half a cell (see stagger()).  Returns false when the reference's 4-point
// scatter would not touch this node at all (so that a non-finite value never spreads further than it does there).
__device__ __forceinline__ bool staggered_weight(int role, double f, double & w) {
    if (role == 0) {
        w = (f >= 0.5) ? 1 - (f - 0.5) : f + 0.5;
        return true;
    }
    if (role > 0) {
        w = 1 - (f + 0.5);
        return f < 0.5;
    }
    w = f - 0.5;
    return f >= 0.5;
}

__device__ __forceinline__ bool node_weight(int role, double f, double & w) {
    w = (role == 0) ? 1 - f : f;
    return role <= 0;
}

// jy_next (origin dy/2 in y) and jz_next (origin dz/2 in z) from the half-step positions (src/System_3d.cpp:472-486,
// staggering :190-193)
__global__ void k_deposit_jy_jz(const Ctl * ctl, Geom g, CellLists L, const double * __restrict__ vjy, const double * __restrict__ vjz,
                                double * __restrict__ jy_next, double * __restrict__ jz_next) {
    QPIC_KERNEL_BEGIN(KID_k_deposit_jy_jz);
    if (cancelled(ctl)) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.n2) return;
    const int j = t / g.nz, k = t - j * g.nz;
    double ajy = 0.0, ajz = 0.0;
    int jr[3], kr[3];
    jr[0] = wrap_dec(j, g.ny);
    jr[1] = j;
    jr[2] = wrap_inc(j, g.ny);
    kr[0] = wrap_dec(k, g.nz);
    kr[1] = k;
    kr[2] = wrap_inc(k, g.nz);
#pragma unroll
    for (int ry = -1; ry <= 1; ry++) {
#pragma unroll
        for (int rz = -1; rz <= 1; rz++) {
            if (ry == 1 && rz == 1) continue;  // contributes to neither array
            const int c = jr[ry + 1] * g.nz + kr[rz + 1];
            const int e = L.start[c + 1];
            for (int s = L.start[c]; s < e; s++) {
                const int p = L.perm ? L.perm[s] : s;
                const double2 f12 = L.fyz[p];
                const double f1 = f12.x, f2 = f12.y;
                double wys, wzn, wyn, wzs;
                // jy_next: staggered in y, node centred in z;  jz_next: node centred in y, staggered in z
                const bool in_jy = staggered_weight(ry, f1, wys) & node_weight(rz, f2, wzn);
                const bool in_jz = node_weight(ry, f1, wyn) & staggered_weight(rz, f2, wzs);
                if (in_jy) ajy += vjy[p] * wys * wzn;
                if (in_jz) ajz += vjz[p] * wyn * wzs;
            }
        }
    }
    jy_next[t] = ajy;
    jz_next[t] = ajz;
}

// ------------------------------------------------------------------------------------------------ potential fix-up
// psi = max(raw - mean(first column U first row), psi_threshold - 1): normalize_potential + increase_minimum
// (src/System_3d.cpp:30-47, :830-841, call sites :364-365 / :468-469).  Every block recomputes the (ny+nz-1)-term
// border mean with the same fixed-order tree, so all blocks subtract bit-identical values without a second launch.
constexpr int FIX_THREADS = 256;

__global__ void __launch_bounds__(FIX_THREADS) k_potential_fixup(const Ctl * ctl, Geom g, const double * __restrict__ raw, double * __restrict__ dst, double floor_value) {
    QPIC_KERNEL_BEGIN(KID_k_potential_fixup);
    if (cancelled(ctl)) return;
    __shared__ double red[FIX_THREADS];
    double part = 0.0;
    const int nb = g.ny + g.nz - 1;
    for (int t = threadIdx.x; t < nb; t += FIX_THREADS) {
        part += (t < g.ny) ? raw[(size_t)g.nz * t] : raw[t - g.ny + 1];
    }
    red[threadIdx.x] = part;
    __syncthreads();
    for (int s = FIX_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    const double border = red[0] / nb;
    for (int t = blockIdx.x * FIX_THREADS + threadIdx.x; t < g.n2; t += gridDim.x * FIX_THREADS) {
        double v = raw[t] - border;
        if (v < floor_value) v = floor_value;
        dst[t] = v;
    }
}

// ------------------------------------------------------------------------------------------------ particle kernels
//
// Two index spaces:
//   p  particle id = the caller's order.  Only the state that survives a slice lives here (y z py pz n, + gamma px for
//      inspection) together with the cell descriptor of (y,z) that the push leaves behind for the cell sort.
//   s  SLOT in the cell-sorted order of the slice-start positions, perm_b[s] = p.  Everything a slice computes per
//      particle is stored by slot, so all M iterations stream contiguous memory and a cell's particles are adjacent
//      (what the tile deposit relies on).  Only two kernels per slice (gamma at the start, push at the end) touch
//      the p-indexed arrays.
struct Particles {
    int np;
    // ---- by particle id
    double *y, *z, *py, *pz, *nw;
    double *gamma, *px;             // src/containers.h:41-42 (kept for inspection / dumps)
    int * cell_p;                   // cell of (y,z)
    int * akey;                     // slab-decomposed sweeps: sort keys of this slice's arrivals (see sort_key); NULL on a single GPU
    double *fy_p, *fz_p;            // fractional position inside it
    // ---- by slot
    int * cell_b;                   // copies of the above in slot order
    // Slot arrays that are always used together are stored as PAIRS (double2): one 16-byte load per lane instead of two 8-byte loads from
    // two streams -- half the load instructions and L1 wavefronts of every kernel below, and a cell's particles read by one thread of a
    // tile kernel become 16-byte pieces of a sector instead of 8-byte ones.
    double2 * fyz_b;                // (fy, fz) of (y, z)
    double *y_s, *z_s, *nw_s;       // (y_s, z_s: only on axes whose step is not a power of two, see slot_position)
    double2 * pyz_s;                // (py, pz) at the slice start
    double *gam, *pxs;              // gamma, px of the current iteration
    double2 * pyzn;                 // (py_next, pz_next)
    int * cell_m;                   // cell of (y_middle, z_middle)
    int * off_m;                    // packed minimal-image cell offset of the half step + half-cell bits (see pack_offset)
    double2 * fyz_m;                // (fy, fz) of the half-step position
    // slice invariants of the B iterations (the reference re-gathers them every iteration, src/System_3d.cpp:419-423,
    // :501-502; (y,z), psi and a_sqr(i) do not change inside the iteration loop)
    double2 * qa;                   // (1 + psi@p, a_sqr@p)
    double2 * t1;                   // dx*0.5*d(a_sqr)/dy@p / (1+psi@p), same for z
    double2 * dp;                   // (dpsi/dy@p, dpsi/dz@p)
    double *vjx, *vrho, *vchi, *vjy, *vjz;  // per-particle deposit values
};

// (oy+2) | (oz+2) << 3 | (fy >= 1/2) << 6 | (fz >= 1/2) << 7 | (kc & 3) << 8
// oy, oz: minimal-image cell offset of the half step; fy, fz: fraction inside the half-step cell; kc: z index of the
// slice-start cell (lets a thread that scans a contiguous run of <= 4 source cells tell which one a particle is from)
__device__ __forceinline__ int pack_offset(int oy, int oz, double fy, double fz, int kc) {
    return (oy + 2) | ((oz + 2) << 3) | ((fy >= 0.5) ? 64 : 0) | ((fz >= 0.5) ? 128 : 0) | ((kc & 3) << 8);
}

// WITH_RHO_CHI = false: inside the B iterations only jx feeds a solve; rho and susceptibility are deposited once per slice by
// a k_tile_deposit<2, true, true> that forms the same two values from gamma, px, n (see there)
template <bool WITH_RHO_CHI>
__device__ __forceinline__ void store_currents(const Particles & P, int s, double nw, double gamma, double px) {
    const double vx = px / gamma;
    const double w = 1 - vx;
    P.vjx[s] = nw * vx / w;
    if (WITH_RHO_CHI) {
        const double r = nw / w;
        P.vrho[s] = r;
        P.vchi[s] = -r / gamma;  // == -nw / (1 - vx) / gamma bit for bit (negation is exact)
    }
}

// fmod(y, l) then +l if negative (src/System_3d.cpp:668-684), without the fmod call in the common cases: for
// 0 <= y < l the result is y; for l <= y < 2l it is y - l, which is exact (Sterbenz), i.e. equal to fmod's; for
// -l < y < 0 fmod returns y and the reference adds l.
__device__ __forceinline__ double wrap_fast(double y, double l) {
    if (y >= 0.0) {
        if (y < l) return y;
        if (y < 2.0 * l) return y - l;
    } else if (y > -l) {
        return y + l;
    }
    return wrap_coordinate(y, l);
}

// Position of a particle from its cell descriptor.  With d a power of two t = y / d is exact, jc = (int)t and f = t - jc are exact, so
// (jc + f) * d IS y bit for bit: the slot copies y_s / z_s (16 bytes per particle and predictor pass of HBM traffic) are not needed on
// such an axis.  Otherwise the stored copy is read.
__device__ __forceinline__ double slot_position(int exact, int jc, double f, double d, const double * __restrict__ stored, int s) {
    return exact ? ((double)jc + f) * d : ld_stream(stored + s);
}

// Canonical order inside a cell = ascending sort key.  The key of an entry is its index in the particle-id arrays -- except for the particles
// that arrived from a neighbouring GPU in this slice: k_migrate_unpack appends them in the order the senders' atomics resolved, so their key is
// (side, slot they had on the sending rank) instead, after every resident: the sums of a slab-decomposed sweep are then bit-reproducible from
// run to run like the single-GPU ones (by induction the senders' slot orders are).  akey == NULL (single GPU): the index itself.
__device__ __forceinline__ long long sort_key(int x, int mig_base1, const int * __restrict__ akey) {
    if (akey == nullptr || mig_base1 == 0 || x < mig_base1 - 1) return x;
    return ((long long)1 << 32) + akey[x - (mig_base1 - 1)];
}

// After the cell sort of the slice-start positions the particle's slice inputs are copied from id order to slot order.  k_sort_rank and that
// copy in one pass (the sort of the slice-start positions is always followed by the slot copies):
// entry t of the bucket-filled list finds its canonical position (ascending particle id inside the cell) and copies its particle's slice
// inputs straight to that slot.  One launch and one pass over perm less per slice.
__global__ void k_rank_to_slots(const Ctl * ctl, Geom g, Particles P, int np, const int * __restrict__ start, const int * __restrict__ tmp, int * __restrict__ perm) {
    QPIC_KERNEL_BEGIN(KID_k_particles_to_slots);
    if (cancelled(ctl)) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= entry_count(ctl, np)) return;
    const int p = tmp[t];
    const int c = P.cell_p[p];
    const int b = start[c], e = start[c + 1];
    // this particle's data: in flight while the bucket is ranked
    const double fy = P.fy_p[p], fz = P.fz_p[p], py = P.py[p], pz = P.pz[p], nw = P.nw[p];
    int rank = 0;
    if (P.akey == nullptr || ctl->mig_base1 == 0) {
        for (int u = b; u < e; u++) rank += (tmp[u] < p) ? 1 : 0;
    } else {
        const int mb = ctl->mig_base1;
        const long long kp = sort_key(p, mb, P.akey);
        for (int u = b; u < e; u++) rank += (sort_key(tmp[u], mb, P.akey) < kp) ? 1 : 0;
    }
    const int s = b + rank;
    perm[s] = p;
    if (s >= live_count(ctl, P.np)) return;  // (slab-decomposed sweeps: the bucket of emigrated entries lies beyond the live slots)
    P.cell_b[s] = c;
    P.fyz_b[s] = make_double2(fy, fz);
    if (!g.exact_dy) P.y_s[s] = P.y[p];
    if (!g.exact_dz) P.z_s[s] = P.z[p];
    P.pyz_s[s] = make_double2(py, pz);
    P.nw_s[s] = nw;
}

// initial gamma, px (src/System_3d.cpp:368-378) + the values deposited at :390-397 + iteration invariants
__global__ void k_particles_gamma(const Ctl * ctl, Geom g, Particles P, const int * __restrict__ perm_b, const double * __restrict__ a_sqr3,
                                  const double * __restrict__ psi, int want_invariants) {
    QPIC_KERNEL_BEGIN(KID_k_particles_gamma);
    if (cancelled(ctl)) return;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= live_count(ctl, P.np)) return;
    const double * a_sqr = a_sqr3 + (size_t)ctl->slice * g.n2;
    const int c = P.cell_b[s];
    const int jc = c / g.nz, kc = c - jc * g.nz;
    const double2 fyz = P.fyz_b[s];
    const double fy = fyz.x, fz = fyz.y;
    const double a_particle = gather_node(a_sqr, g.ny, g.nz, jc, kc, fy, fz);
    const double psi_particle = gather_node(psi, g.ny, g.nz, jc, kc, fy, fz);
    const double2 pyz = P.pyz_s[s];
    const double py = pyz.x, pz = pyz.y, nw = P.nw_s[s];
    const double q = 1 + psi_particle;
    // Divisions by 1 + psi@p are products with ONE reciprocal (within an ulp of the reference's quotients; tolerance 1e-12): the four
    // quotients of this kernel and the six of the predictor were ~60 % of their instructions.
    const double rq = 1.0 / q;
    const double gamma = 0.5 * (1 + py * py + pz * pz + a_particle + q * q) * rq;
    const double px = 0.5 * (1 + py * py + pz * pz + a_particle - q * q) * rq;
    if (!want_invariants) {
        // no B iterations / no push follow (M = 0 or slice 0): this is the final gamma, px of the particle (the push
        // writes them otherwise)
        const int p = perm_b[s];
        P.gamma[p] = gamma;
        P.px[p] = px;
    }
    P.gam[s] = gamma;
    if (!want_invariants) {
        // only when no B iteration follows are these read again (deposits at :390-397, the slice-end deposit); otherwise every
        // predictor pass overwrites them before their first use
        P.pxs[s] = px;
        store_currents<true>(P, s, nw, gamma, px);
    }
    if (want_invariants) {
        P.qa[s] = make_double2(q, a_particle);
        P.t1[s] = make_double2(g.dx * 0.5 * gather_yder(a_sqr, g.ny, g.nz, g.inv_dy, jc, kc, fy, fz) * rq,
                               g.dx * 0.5 * gather_zder(a_sqr, g.ny, g.nz, g.inv_dz, jc, kc, fy, fz) * rq);
        P.dp[s] = make_double2(gather_yder(psi, g.ny, g.nz, g.inv_dy, jc, kc, fy, fz), gather_zder(psi, g.ny, g.nz, g.inv_dz, jc, kc, fy, fz));
    }
}

// What a predictor pass needs per particle besides py, pz and the B gathers: the slice invariants and the previous iteration's gamma.
// FIRST (iteration 0 of a slice): k_particles_gamma's work is done here instead -- psi, a_sqr and their derivatives are gathered at the particle,
// the initial gamma (src/System_3d.cpp:368-378) is formed, and predict_store_invariants leaves the invariants for iterations 1 .. M-1.  That saves
// the separate pass over the slot arrays (100 bytes per particle, one launch) that only produced inputs of this kernel.  Later iterations stream
// the stored values (read once per pass: evict-first).
struct PredictIn {
    double q, gamma, a_particle, t1y, t1z, dpy, dpz;
};

template <bool FIRST>
__device__ __forceinline__ PredictIn predict_inputs(const Ctl * ctl, const Geom & g, const Particles & P, int s, int jc, int kc, double fy, double fz, double py, double pz,
                                                    const double * __restrict__ a_sqr3, const double * __restrict__ psi) {
    PredictIn in;
    if (FIRST) {
        const double * a_sqr = a_sqr3 + (size_t)ctl->slice * g.n2;
        in.a_particle = gather_node(a_sqr, g.ny, g.nz, jc, kc, fy, fz);
        const double psi_particle = gather_node(psi, g.ny, g.nz, jc, kc, fy, fz);
        in.q = 1 + psi_particle;
        const double rq = 1.0 / in.q;  // (see k_particles_gamma)
        in.gamma = 0.5 * (1 + py * py + pz * pz + in.a_particle + in.q * in.q) * rq;
        in.t1y = g.dx * 0.5 * gather_yder(a_sqr, g.ny, g.nz, g.inv_dy, jc, kc, fy, fz) * rq;
        in.t1z = g.dx * 0.5 * gather_zder(a_sqr, g.ny, g.nz, g.inv_dz, jc, kc, fy, fz) * rq;
        in.dpy = gather_yder(psi, g.ny, g.nz, g.inv_dy, jc, kc, fy, fz);
        in.dpz = gather_zder(psi, g.ny, g.nz, g.inv_dz, jc, kc, fy, fz);
    } else {
        const double2 qa = ld_stream2(P.qa + s), t1 = ld_stream2(P.t1 + s), dp = ld_stream2(P.dp + s);
        in.q = qa.x;
        in.a_particle = qa.y;
        in.gamma = P.gam[s];
        in.t1y = t1.x, in.t1z = t1.y, in.dpy = dp.x, in.dpz = dp.y;
    }
    return in;
}

template <bool FIRST>
__device__ __forceinline__ void predict_store_invariants(const Particles & P, int s, const PredictIn & in) {
    if (FIRST) {
        P.qa[s] = make_double2(in.q, in.a_particle);
        P.t1[s] = make_double2(in.t1y, in.t1z);
        P.dp[s] = make_double2(in.dpy, in.dpz);
    }
}

// ---- far movers: half steps beyond the +-1 cell window of the tile deposits (round 2)
// The tile kernels find the particles that land in a target cell by looking at the cells around it, which is only complete while every half step
// stays within one cell.  A particle that goes further (bubble sheath, trajectory crossing) used to raise Ctl::fast_fail: the host discarded the slice
// and redid it -- and the next 16, 32, ... -- on the sorted path.  Now the predictor appends such a particle to a per-TARGET-cell list in global memory
// (integer atomics: only the order of a list depends on scheduling) and marks its offset code FAR; the psi_middle source deposit merges a target
// cell's far entries into its shared-memory mover list before the list is sorted by slot, so they are accumulated -- and published for the jy/jz deposit
// -- exactly where ascending slot order puts them: same sums, bit for bit, as the sorted path.  The counters are double buffered over the B
// iterations (halo target cells are read by two blocks, so the reader cannot clear them): deposit `it` reads cnt[it & 1] and clears cnt[(it + 1) & 1];
// the slice's first charge deposit clears cnt[0].  Only a list overflow (more than FAR_K far arrivals in one cell, or a crowded cell that also has far
// arrivals) still takes the rollback.
constexpr int FAR_K = 8;
constexpr int FAR_CODE = 7 | (7 << 3);  // bits 0-5 of an offset code: neither a stayer nor inside any tile window
struct FarMovers {
    int * cnt;   // [n2] arrivals per target cell of THIS iteration (NULL: far movers abort, as in round 1)
    int * lst;   // [FAR_K][n2] entries slot << 2 | hy << 1 | hz
    int * clear; // [n2] the other buffer: cleared by the deposit's owner threads
    // slab-decomposed sweeps: the target rows a rank can deposit into = its own rows and the halo rows whose partial sums the halo reduce
    // carries to the neighbour (row_n == 0: every row, single GPU); a far mover that lands outside still takes the abort
    int row_lo, row_n;
};

__device__ __forceinline__ bool far_row_ok(int row_lo, int row_n, int jm, int ny) {
    if (row_n == 0) return true;
    int rel = jm - row_lo;
    rel += (rel < 0) ? ny : 0;
    rel -= (rel >= ny) ? ny : 0;
    return rel < row_n;
}

// minimal-image cell offset jm - jc folded into [-n/2, n/2]
__device__ __forceinline__ int cell_offset(int jm, int jc, int n) {
    int o = jm - jc;
    if (2 * o > n) o -= n;
    if (2 * o < -n) o += n;
    return o;
}

// predictor: py_next, pz_next (src/System_3d.cpp:416-433), half-step position (:436-443) and its cell descriptor.
// FAST (sort-free deposit): records the cell offset of the half step instead of counting cell populations, and raises
// Ctl::fast_fail if a particle leaves the +-1 cell window the tile deposit covers.
// The corrector gamma / px (src/System_3d.cpp:498-508) and the values deposited at :520-527 depend only on py_next, pz_next,
// so they are formed here as well: one streaming pass over the slot arrays per iteration instead of two.
// store_px: only the last B iteration's px is read again (rho / susceptibility deposit, push); the general path's deposits read it
// every iteration.
#ifndef QPIC_PREDICT_MINB
#define QPIC_PREDICT_MINB 1
#endif
template <bool FAST, bool FIRST = false>
__global__ void __launch_bounds__(256, QPIC_PREDICT_MINB) k_particles_predict(Ctl * ctl, Geom g, Particles P, const double * __restrict__ by, const double * __restrict__ bz,
                                    int * __restrict__ count, int store_px, FarMovers far, const double * __restrict__ a_sqr3, const double * __restrict__ psi) {
    QPIC_KERNEL_BEGIN(KID_k_particles_predict);
    // The control block is tested only before the first store: its load then overlaps the streaming loads instead of
    // heading the dependent chain (every slot < P.np is readable whatever the flags say).
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.np) return;
    const int c = P.cell_b[s];
    const int jc = c / g.nz, kc = c - jc * g.nz;
    const double2 fyz = P.fyz_b[s];
    const double fy = fyz.x, fz = fyz.y;
    // issue the independent streaming loads (16 bytes each) before the dependent field gathers
    const double2 pyz = ld_stream2(P.pyz_s + s);
    const double py = pyz.x, pz = pyz.y;
    const PredictIn in = predict_inputs<FIRST>(ctl, g, P, s, jc, kc, fy, fz, py, pz, a_sqr3, psi);
    const double q = in.q, gamma = in.gamma, a_particle = in.a_particle;
    const double t1y = in.t1y, t1z = in.t1z, dpy = in.dpy, dpz = in.dpz;
    const double y0 = slot_position(g.exact_dy, jc, fy, g.dy, P.y_s, s), z0 = slot_position(g.exact_dz, kc, fz, g.dz, P.z_s, s);
    const double nw = P.nw_s[s];
    int js, ks;
    double fys, fzs;
    stagger(jc, fy, g.ny, js, fys);
    stagger(kc, fz, g.nz, ks, fzs);
    // by lives at (y, z + dz/2), bz at (y + dy/2, z)  (src/System_3d.cpp:180-181)
    const double by_particle = gather4(by, g.nz, jc, wrap_inc(jc, g.ny), ks, wrap_inc(ks, g.nz), fy, fzs);
    const double bz_particle = gather4(bz, g.nz, js, wrap_inc(js, g.ny), kc, wrap_inc(kc, g.nz), fys, fz);
    if (cancelled(ctl) || s >= live_count(ctl, P.np)) return;
    QPIC_KERNEL_TAIL();
    predict_store_invariants<FIRST>(P, s, in);
    const double rq = 1.0 / q;  // (see k_particles_gamma)
    double pyn = py - t1y;
    pyn += g.dx * gamma * dpy * rq;
    pyn -= g.dx * bz_particle;
    double pzn = pz - t1z;
    pzn += g.dx * gamma * dpz * rq;
    pzn += g.dx * by_particle;
    P.pyzn[s] = make_double2(pyn, pzn);
    {
        // corrector (the gamma read above was the previous iteration's; this thread owns element s)
        const double p_squared = 0.25 * (py + pyn) * (py + pyn) + 0.25 * (pz + pzn) * (pz + pzn);
        const double gamma_c = 0.5 * (1 + p_squared + a_particle + q * q) * rq;
        const double px_c = 0.5 * (1 + p_squared + a_particle - q * q) * rq;
        P.gam[s] = gamma_c;
        if (store_px) P.pxs[s] = px_c;
        store_currents<!FAST>(P, s, nw, gamma_c, px_c);  // the general path's node-centric deposit reads vrho / vchi
    }
    const double ym = wrap_fast(y0 + 0.5 * g.dx * pyn * rq, g.ly);
    const double zm = wrap_fast(z0 + 0.5 * g.dx * pzn * rq, g.lz);
    int jm, km;
    double f1, f2;
    locate_fast(ym, g.dy, g.inv_dy, g.exact_dy, g.ny, jm, f1);
    locate_fast(zm, g.dz, g.inv_dz, g.exact_dz, g.nz, km, f2);
    const int cm = jm * g.nz + km;
    if (!FAST) P.cell_m[s] = cm;  // only the general path's cell sort reads it
    P.fyz_m[s] = make_double2(f1, f2);
    if (FAST) {
        const int oy = cell_offset(jm, jc, g.ny), oz = cell_offset(km, kc, g.nz);
        // !(f >= 0) catches the NaN positions of a blown-up run
        if (!(f1 >= 0.0) || !(f2 >= 0.0)) {
            atomicExch(&ctl->fast_fail, 1);
            P.off_m[s] = pack_offset(0, 0, 0.0, 0.0, kc);
        } else if (oy < -1 || oy > 1 || oz < -1 || oz > 1) {
            // far mover: listed under its target cell (see FarMovers)
            const int hy = f1 >= 0.5 ? 1 : 0, hz = f2 >= 0.5 ? 1 : 0;
            const int pos = (far.cnt && far_row_ok(far.row_lo, far.row_n, jm, g.ny)) ? atomicAdd(&far.cnt[cm], 1) : FAR_K;
            if (pos < FAR_K) far.lst[(size_t)pos * g.n2 + cm] = (s << 2) | (hy << 1) | hz;
            else atomicExch(&ctl->fast_fail, 1);
            P.off_m[s] = FAR_CODE | (hy << 6) | (hz << 7) | ((kc & 3) << 8);
        } else {
            P.off_m[s] = pack_offset(oy, oz, f1, f2, kc);
        }
    } else {
        atomicAdd(&count[cm], 1);
    }
}

// values for the jy/jz deposit: n py_next / (1 + psi_middle@(y_middle, z_middle))  (src/System_3d.cpp:480-486)
__global__ void k_particles_currents_yz(const Ctl * ctl, Geom g, Particles P, const double * __restrict__ psi_middle) {
    QPIC_KERNEL_BEGIN(KID_k_particles_currents_yz);
    if (cancelled(ctl)) return;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= live_count(ctl, P.np)) return;
    const int cm = P.cell_m[s];
    const int jm = cm / g.nz, km = cm - jm * g.nz;
    const double2 fm = P.fyz_m[s];
    const double psi_particle = gather_node(psi_middle, g.ny, g.nz, jm, km, fm.x, fm.y);
    const double nw = P.nw_s[s];
    const double w = 1 + psi_particle;
    const double2 pn = P.pyzn[s];
    P.vjy[s] = nw * pn.x / w;
    P.vjz[s] = nw * pn.y / w;
}

// ---- particle migration of slab-decomposed runs (qpic_dist.cuh): where emigrants of k_particles_advance go
struct MigrateView {
    int enabled;
    int row0, ny_loc, cap;
    double * out[2];   // peer inboxes: [0] of the rank below (its "from above" box), [1] of the rank above; SoA [6][cap] (y z py pz n + sender's slot)
    int * out_count;   // local int[2]
    int * error;
};

// py <- py_next, pz <- pz_next (src/System_3d.cpp:593-598), coordinate advance with psi_middle gathered at (y,z)
// (:601-608), new cell descriptor + population count for the next slice's lists; back to particle-id order.
// Slab-decomposed runs: the id of a particle is simply its slot (perm_b == NULL); a particle whose new cell belongs to a
// neighbouring rank is appended to that rank's inbox and its own entry is parked in the extra "emigrated" bucket n2.
__global__ void k_particles_advance(const Ctl * ctl, Geom g, Particles P, const int * __restrict__ perm_b, const double * __restrict__ psi_middle,
                                    int * __restrict__ count, MigrateView mig, int track_gamma_px) {
    QPIC_KERNEL_BEGIN(KID_k_particles_advance);
    if (cancelled(ctl)) return;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= live_count(ctl, P.np)) return;
    const int p = perm_b ? perm_b[s] : s;
    const int c = P.cell_b[s];
    const int jc = c / g.nz, kc = c - jc * g.nz;
    const double2 fyz = P.fyz_b[s];
    const double psi_particle = gather_node(psi_middle, g.ny, g.nz, jc, kc, fyz.x, fyz.y);
    const double2 pn = P.pyzn[s];
    const double py = pn.x, pz = pn.y;
    P.py[p] = py;
    P.pz[p] = pz;
    if (track_gamma_px || ctl->slice <= 1) {
        // gamma, px by particle id exist for inspection only (qpic_download_particles): slice 0 rewrites them (k_particles_gamma), so
        // only the push of slice 1 has to leave them behind -- unless the caller wants them after every slice (QPIC_TRACK_GAMMA_PX=1)
        P.gamma[p] = P.gam[s];
        P.px[p] = P.pxs[s];
    }
    const double rw = 1.0 / (1 + psi_particle);  // (see k_particles_gamma)
    const double y = wrap_fast(slot_position(g.exact_dy, jc, fyz.x, g.dy, P.y_s, s) + g.dx * py * rw, g.ly);
    const double z = wrap_fast(slot_position(g.exact_dz, kc, fyz.y, g.dz, P.z_s, s) + g.dx * pz * rw, g.lz);
    P.y[p] = y;
    P.z[p] = z;
    int jn, kn;
    double f1, f2;
    locate_fast(y, g.dy, g.inv_dy, g.exact_dy, g.ny, jn, f1);
    locate_fast(z, g.dz, g.inv_dz, g.exact_dz, g.nz, kn, f2);
    int cn = jn * g.nz + kn;
    if (mig.enabled) {
        const double nw = P.nw_s[s];
        P.nw[p] = nw;
        int rel = jn - mig.row0;  // row relative to the slab, minimal image
        if (2 * rel > g.ny) rel -= g.ny;
        if (2 * rel < -g.ny) rel += g.ny;
        if (rel < 0 || rel >= mig.ny_loc) {
            const int side = rel < 0 ? 0 : 1;
            if (rel < -mig.ny_loc || rel >= 2 * mig.ny_loc) atomicExch(mig.error, 4);  // beyond the neighbouring slab
            const int i = atomicAdd(&mig.out_count[side], 1);
            if (i < mig.cap) {
                double * b = side == 0 ? mig.out[0] : mig.out[1];
                b[0 * (size_t)mig.cap + i] = y;
                b[1 * (size_t)mig.cap + i] = z;
                b[2 * (size_t)mig.cap + i] = py;
                b[3 * (size_t)mig.cap + i] = pz;
                b[4 * (size_t)mig.cap + i] = nw;
                b[5 * (size_t)mig.cap + i] = (double)s;  // sort key of the arrival on the receiving rank (see sort_key)
            }
            cn = g.n2;  // the "emigrated" bucket
        }
    }
    P.cell_p[p] = cn;
    P.fy_p[p] = f1;
    P.fz_p[p] = f2;
    atomicAdd(&count[cn], 1);
}

// ------------------------------------------------------------------------------------------------ end of slice
// output_step (src/System_3d.cpp:629-666): ex(i), ey(i), ez(i); by(i)/bz(i) go to their 3-D rows; optional per-slice
// copies of psi,rho,jx,jy,jz (write_slice, :663-665); then the K1/K31 copies psi_prev<-psi, jy<-jy_next, jz<-jz_next
// (:334-339, :612-618) which are elementwise and therefore safe to do here.
struct SliceOut {
    double *ex3, *ey3, *ez3, *by3, *bz3;
    double *psi3, *rho3, *jx3, *jy3, *jz3;  // nullable
};

__global__ void k_end_of_slice(const Ctl * ctl, Geom g, SliceOut o, const double * __restrict__ psi, double * __restrict__ psi_prev,
                               const double * __restrict__ by, const double * __restrict__ bz, const double * __restrict__ rho,
                               const double * __restrict__ jx, double * __restrict__ jy, double * __restrict__ jz,
                               const double * __restrict__ jy_next, const double * __restrict__ jz_next, int b_is_zero, int roll, int ignore_stop,
                               int row0, int nrows, Ctl * ctl_rw, int * __restrict__ ticket) {
    QPIC_KERNEL_BEGIN(KID_k_end_of_slice);
    if (cancelled(ctl) && !ignore_stop) return;
    // ticket != NULL (slice body): the block that finishes last moves the sweep to the next slice (every block has read
    // ctl->slice by then; nothing that runs concurrently with this kernel reads it)
    struct Advance {
        Ctl * ctl;
        int * ticket;
        __device__ ~Advance() {
            if (ticket == nullptr || threadIdx.x != 0) return;
            __threadfence();
            if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
                *ticket = 0;
                ctl->slice -= 1;
            }
        }
    };
    const int i = ignore_stop ? 0 : ctl->slice;
    __syncthreads();  // the whole block has read the control block before thread 0 may leave
    Advance advance_on_exit{ctl_rw, ticket};
    // rows [row0, row0 + nrows): all of them on a single GPU, the own slab in a distributed sweep
    if (blockIdx.x * blockDim.x + threadIdx.x >= (unsigned)(nrows * g.nz)) return;
    const int t = row0 * g.nz + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = t / g.nz, k = t - j * g.nz;
    const size_t off = (size_t)i * g.n2 + t;
    const double ps = psi[t];
    if (i < g.nx - 1) o.ex3[off] = (ps - psi_prev[t]) * g.inv_dx;
    const double b_y = b_is_zero ? 0.0 : by[t];
    const double b_z = b_is_zero ? 0.0 : bz[t];
    o.ey3[off] = -(psi[wrap_inc(j, g.ny) * g.nz + k] - ps) * g.inv_dy + b_z;
    o.ez3[off] = -(psi[j * g.nz + wrap_inc(k, g.nz)] - ps) * g.inv_dz - b_y;
    if (!b_is_zero) {
        o.by3[off] = b_y;
        o.bz3[off] = b_z;
    }
    if (o.psi3) {
        o.psi3[off] = ps;
        o.rho3[off] = rho[t];
        o.jx3[off] = jx[t];
        o.jy3[off] = jy[t];
        o.jz3[off] = jz[t];
    }
    if (roll) {
        psi_prev[t] = ps;
        jy[t] = jy_next[t];
        jz[t] = jz_next[t];
    }
}

// ---- Fields_xy output (Output_writer with output_xy, src/Output_writer.cpp:176-187): cuts at z = z0, linear in z between the
// two neighbouring planes of the array's own z-origin (calculate_xy_slice / calculate_y_slice, src/array_utils.cpp:7-45).
// k1 = k0 + 1 wraps at the box edge (the reference would read past the row there).
struct XyCut {
    int k0, k1;
    double frac;
};

// cut of `count` slices of a 3-D array starting at a3 -> dst[count][ny]
__global__ void k_xy_cut(Geom g, const double * __restrict__ a3, int count, XyCut c, double * __restrict__ dst) {
    QPIC_KERNEL_BEGIN(KID_k_xy_cut);
    const size_t total = (size_t)count * g.ny;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const double * row = a3 + t * g.nz;
        dst[t] = (1 - c.frac) * row[c.k0] + c.frac * row[c.k1];
    }
}

// the per-slice datasets psi, rho, jx, jy, jz (write_slice, src/System_3d.cpp:663-665): line i of their (nx, ny) cuts, taken
// BEFORE the jy / jz roll like the slices themselves (:610 vs :612-618).  node: z-origin 0; half: z-origin dz/2 (jz)
struct XyLines {
    double *psi, *rho, *jx, *jy, *jz;  // (nx, ny) each
};

__global__ void k_xy_lines(const Ctl * ctl, Geom g, XyLines o, XyCut node, XyCut half, const double * __restrict__ psi, const double * __restrict__ rho,
                           const double * __restrict__ jx, const double * __restrict__ jy, const double * __restrict__ jz, int ignore_stop, int row0,
                           int nrows) {
    QPIC_KERNEL_BEGIN(KID_k_xy_lines);
    if (cancelled(ctl) && !ignore_stop) return;
    const int jl = blockIdx.x * blockDim.x + threadIdx.x;
    if (jl >= nrows) return;
    const int j = row0 + jl;
    const int i = ignore_stop ? 0 : ctl->slice;
    const size_t at = (size_t)i * g.ny + j;
    const size_t r = (size_t)j * g.nz;
    o.psi[at] = (1 - node.frac) * psi[r + node.k0] + node.frac * psi[r + node.k1];
    o.rho[at] = (1 - node.frac) * rho[r + node.k0] + node.frac * rho[r + node.k1];
    o.jx[at] = (1 - node.frac) * jx[r + node.k0] + node.frac * jx[r + node.k1];
    o.jy[at] = (1 - node.frac) * jy[r + node.k0] + node.frac * jy[r + node.k1];
    o.jz[at] = (1 - half.frac) * jz[r + half.k0] + half.frac * jz[r + half.k1];
}


// ------------------------------------------------------------------------------------------------ sort-free tile deposits
// Deterministic deposition without sorting the half-step positions globally.  A block owns a TILE x TILE patch of grid
// nodes = the TILE_ACC x TILE_ACC target cells whose CIC corners touch them (one halo row / column below the patch, recomputed
// identically by the neighbouring block).  The slot order groups particles by their cell at the start of the slice, so a
// cell's particles are one contiguous slot range.  Every target cell sums the four CIC corner contributions
// (src/array_utils.cpp:75-82) of the particles that land in it IN ASCENDING SLOT ORDER into four private corner sums; after a
// barrier every node adds the corner sums of its four neighbouring target cells in a fixed order.  No floating-point
// atomics, no dependence on scheduling: bit-reproducible.
//
//   k_tile_deposit       deposits at (y,z): the target cell IS the source cell, a thread walks its own cell's particles.
//   half-step deposits   (k_tile_stay_deposit / k_tile_assemble for the psi_middle source, k_tile_cell_deposit for jy / jz): ~98 % of the particles
//                        end their half step in the cell they started in and are summed in place; the movers are binned per TARGET cell
//                        (integer atomics -- only the order of a list depends on scheduling), the target cell sorts its short list by slot
//                        (canonical order) and accumulates movers from lower cells, its stayers, movers from higher cells.  Lists longer
//                        than BIN_K fall back to an in-order scan of the source window.  (Rounds 1-2 went through three earlier
//                        generations -- presence-mask window scans, k_tile_bin_deposit, k_tile_list_deposit -- that are documented in
//                        DESIGN.md section 4.1 and live in the commit history.)
constexpr int TILE = 16;
constexpr int TILE_ACC = TILE + 1;
constexpr int TILE_CELLS = TILE_ACC * TILE_ACC;
constexpr int TILE_MAX_OUT = 3;
#ifndef QPIC_BIN_K
#define QPIC_BIN_K 16  // list capacity per target cell and output (tests build the emulator with 1 to force the in-order fallback)
#endif
constexpr int BIN_K = QPIC_BIN_K;

struct TileOutput {
    const double * val;    // per-particle value (slot order)
    double sign;           // +1 / -1 (charge deposits use -n)
    int bg_mode;           // 0: none, 1: jx_bunch(i) - rho_bunch(i) - rho_ion (psi source, :345), 2: bg3(i)
    const double * bg3;    // bg_mode 2: a 3-D array whose slice i initialises the sum (rho_bunch / jx_bunch, :384-385)
    double * out;          // 2-D array, or a 3-D array when out_3d
    int out_3d;
};

struct TileJob {
    int nout;
    // slab-decomposed runs launch only the tile rows that the rank's particles can reach (own rows + one tile row on either
    // side) and add the background terms on own rows only; single GPU: tile_row0 = 0, tile_rows = all, own = [0, ny)
    int tile_row0, tile_rows;
    int own_row0, own_rows;
    TileOutput o[TILE_MAX_OUT];
    const double *jxb3, *rhob3, *rho_ion;  // for bg_mode 1
    const double *psi_mid, *nw;            // jy/jz job
    // Per-cell MOVER lists of the half-step positions, left in global memory by the psi_middle source kernel (k_tile_stay_deposit /
    // k_tile_assemble) for the jy/jz deposit of the same iteration (k_tile_cell_deposit): bin_cnt[cell] entries bin_lst[i * n2 + cell], sorted by
    // slot, entry = slot << 2 | hy << 1 | hz; a count above BIN_K means "not listed, scan the window".  list_row0 / list_rows:
    // the cell rows whose lists were written (all of them on a single GPU).
    int * bin_cnt;
    int * bin_lst;
    int list_row0, list_rows;
    int cell_row0, cell_rows;  // k_tile_cell_deposit: node rows launched (0, ny on a single GPU)
    int rc_folded;             // k_tile_deposit<2, true, true>: last B iteration + 1 when k_tile_fold_rho_chi serves the slices whose last predictor pass ran (0: off)
    int * clear_cnt;           // experimental predictor-side deposit: mover counters cleared by the slice's first charge deposit (per node)
    // jy/jz job on a single GPU: the node that has just formed jy_next, jz_next also forms the sources of the By + i Bz solve
    // (src/System_3d.cpp:489-495, :530-536, :555-564 incl. the :563 quirk) into b_src, so that the forward rows of that solve
    // load one complex array instead of seven real ones with their stencils.  NULL: the row kernel builds them.
    double2 * b_src;
    const double *b_by, *b_bz, *b_jx, *b_jy, *b_jz;
    double b_D;
};

__device__ __forceinline__ int mod_wrap(int j, int n) {
    j %= n;
    return j < 0 ? j + n : j;
}

// the same for -n <= j < 2n (tile coordinates with their halo), without the integer division
__device__ __forceinline__ int wrap_near(int j, int n) {
    j += (j < 0) ? n : 0;
    j -= (j >= n) ? n : 0;
    return j;
}

// nodes of the patch: own cell's (1-fy)(1-fz) sum, then the y-, z- and diagonal neighbours' sums, fixed order.
// acc layout: [output][corner][target cell] (conflict-free), target cell (ty, tz) = (node - patch origin + 1)
// rho_ion at the node of target cell (ty, tz), fetched at the start of a kernel whose first output has bg_mode 1 so that the
// load is not part of the dependent chain at its end
template <int NOUT>
__device__ __forceinline__ double tile_prefetch_ion(const Geom & g, const TileJob & job, int ty, int tz, int j, int k) {
    if (job.o[0].bg_mode != 1 || ty < 1 || tz < 1 || j >= g.ny || k >= g.nz) return 0.0;
    return job.rho_ion[j * g.nz + k];
}

template <int NOUT>
__device__ __forceinline__ void tile_store_nodes(const Ctl * ctl, const Geom & g, const TileJob & job, const double * acc, int ty, int tz, int j, int k,
                                                 double ion) {
    if (ty < 1 || tz < 1 || j >= g.ny || k >= g.nz) return;
    const int t = j * g.nz + k;
    const size_t off3 = (size_t)ctl->slice * g.n2 + t;
    const bool own = (unsigned)(j - job.own_row0) < (unsigned)job.own_rows;
    if (job.clear_cnt) job.clear_cnt[t] = 0;
#pragma unroll
    for (int a = 0; a < NOUT; a++) {
        const TileOutput & o = job.o[a];
        double v = 0.0;
        if (o.bg_mode == 1 && own) {
            const double jb_ = job.jxb3 ? job.jxb3[off3] : 0.0;
            const double rb_ = job.rhob3 ? job.rhob3[off3] : 0.0;
            v = jb_ - rb_ - (a == 0 ? ion : job.rho_ion[t]);
        } else if (o.bg_mode == 2 && o.bg3 && own) {
            v = o.bg3[off3];
        }
        const double * acc_a = acc + a * 4 * TILE_CELLS;
        v += acc_a[0 * TILE_CELLS + ty * TILE_ACC + tz];
        v += acc_a[1 * TILE_CELLS + (ty - 1) * TILE_ACC + tz];
        v += acc_a[2 * TILE_CELLS + ty * TILE_ACC + tz - 1];
        v += acc_a[3 * TILE_CELLS + (ty - 1) * TILE_ACC + tz - 1];
        o.out[o.out_3d ? off3 : (size_t)t] = v;
    }
}

// RHO_CHI (NOUT = 2): the outputs are rho and susceptibility(i) and the deposited values n / (1 - px/gamma), -n / (1 - px/gamma) / gamma
// (src/System_3d.cpp:520-527, same expressions as store_currents) are formed here from the corrector's gamma, px.
#ifndef QPIC_TD_MINB
#define QPIC_TD_MINB 0
#endif
template <int NOUT, bool FINAL_PASS, bool RHO_CHI = false>
__global__ void
#if QPIC_TD_MINB > 0
    __launch_bounds__(TILE_CELLS, QPIC_TD_MINB)
#else
    __launch_bounds__(TILE_CELLS)
#endif
    k_tile_deposit(const Ctl * ctl, Geom g, const int * __restrict__ start, Particles P, TileJob job) {
    QPIC_KERNEL_BEGIN(KID_COUNT + (job.o[0].bg_mode == 2 ? 2 : 0) + (FINAL_PASS ? 8 : 0));
    // (the control block is tested after the particle loop: see k_particles_predict)
    __shared__ double acc[NOUT * 4 * TILE_CELLS];
    const int ay = threadIdx.y, az = threadIdx.x;
    const int tcell = ay * TILE_ACC + az;
    const int jt = ((job.tile_row0 + (int)blockIdx.y) % job.tile_rows) * TILE - 1 + ay, kt = blockIdx.x * TILE - 1 + az;  // unwrapped target cell
    const int cell = wrap_near(jt, g.ny) * g.nz + wrap_near(kt, g.nz);
    if (RHO_CHI && job.rc_folded) {
        // only the slices that the |B| stop condition ended before their last predictor pass are left to this kernel (see k_tile_fold_rho_chi)
        const int tag = ctl->stop_tag;
        if (tag == 0 || (tag - 1) % STOP_TAG_STRIDE == job.rc_folded - 1) return;
    }
    const int s0 = start[cell], s1 = start[cell + 1];
    const double ion = tile_prefetch_ion<NOUT>(g, job, ay, az, jt, kt);
    double c00[NOUT], c10[NOUT], c01[NOUT], c11[NOUT];
#pragma unroll
    for (int a = 0; a < NOUT; a++) c00[a] = c10[a] = c01[a] = c11[a] = 0.0;
    const bool stopped_early = RHO_CHI && ctl->stop_tag != 0;
    // chunks of 4 particles: all loads of a chunk are issued before the (ordered) accumulation
    for (int sb = s0; sb < s1; sb += 4) {
        double fy[4], fz[4], v[NOUT][4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int s = sb + u < s1 ? sb + u : s0;
            const double2 fyz = P.fyz_b[s];
            fy[u] = fyz.x;
            fz[u] = fyz.y;
            if (RHO_CHI) {
                const double gamma = P.gam[s];
                // px of the last executed iteration.  The predictor stores it in its last pass only; when the |B| stop condition
                // ended the slice in an earlier pass (this deposit then still runs: rho / susceptibility of the stopping iteration
                // are output, src/System_3d.cpp:520-527 before :545-550) it is rebuilt from gamma - px = 1 + psi, exact up to rounding.
                const double px = stopped_early ? gamma - P.qa[s].x : P.pxs[s];
                const double vx = px / gamma;
                const double r = P.nw_s[s] / (1 - vx);
                v[0][u] = r;
                v[NOUT - 1][u] = -r / gamma;
            } else {
#pragma unroll
                for (int a = 0; a < NOUT; a++) v[a][u] = job.o[a].val[s];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (sb + u < s1) {
#pragma unroll
                for (int a = 0; a < NOUT; a++) {
                    const double vv = job.o[a].sign * v[a][u];
                    c00[a] += vv * (1 - fy[u]) * (1 - fz[u]);
                    c10[a] += vv * fy[u] * (1 - fz[u]);
                    c01[a] += vv * (1 - fy[u]) * fz[u];
                    c11[a] += vv * fy[u] * fz[u];
                }
            }
        }
    }
    if (FINAL_PASS) {
        if (ctl->fast_fail || (ctl->stop_tag != 0 && (ctl->stop_tag - 1) / STOP_TAG_STRIDE != ctl->slice)) return;
    } else if (cancelled(ctl)) {
        return;
    }
    QPIC_KERNEL_TAIL();
#pragma unroll
    for (int a = 0; a < NOUT; a++) {
        double * acc_a = acc + a * 4 * TILE_CELLS;
        acc_a[0 * TILE_CELLS + tcell] = c00[a];
        acc_a[1 * TILE_CELLS + tcell] = c10[a];
        acc_a[2 * TILE_CELLS + tcell] = c01[a];
        acc_a[3 * TILE_CELLS + tcell] = c11[a];
    }
    __syncthreads();
    tile_store_nodes<NOUT>(ctl, g, job, acc, ay, az, jt, kt, ion);
}

// ------------------------------------------------------------------------------------------------ psi_middle source, stayers direct
// Successor of the shared-memory binning deposit of round 1 (k_tile_bin_deposit<0>, removed).  In laminar flow ~98 % of the particles end their half step in
// the cell they started in ("stayers": offset code 0, 0).  Those need no binning: thread (sy, sz), which is the source-cell thread
// AND the target-cell thread of cell (j0 - 2 + sy, k0 - 2 + sz), keeps its own first four particles (codes, fractions, weights) in
// registers -- all loaded in ONE round trip after the slot range, before the barrier -- and only the movers go through the
// shared-memory lists.  After the barrier a target cell accumulates in ascending slot order as before: movers that arrived from
// lower cells, then its stayers (from registers), then movers from higher cells.  Same sums bit for bit as the binning kernel's;
// the dependent chain of a block shrinks from [range -> codes -> barrier -> entry -> data ...] to [range -> codes + data -> barrier].
// Published for the jy/jz deposit of the iteration (k_tile_cell_deposit): the MOVERS of every cell only (count + sorted entries);
// that kernel finds the stayers through the slot range and the offset codes, like this one.
constexpr int STAY_CODE = 2 | (2 << 3);  // pack_offset's bits 0-5 for oy = oz = 0

inline size_t tile_stay_smem_bytes() { return (size_t)TILE_CELLS * (4 * sizeof(double) + (BIN_K + 1) * sizeof(int)); }

#ifndef QPIC_STAY_MINB
#define QPIC_STAY_MINB 1
#endif
__global__ void __launch_bounds__((TILE_ACC + 2) * (TILE_ACC + 2), QPIC_STAY_MINB)
    k_tile_stay_deposit(Ctl * ctl, Geom g, const int * __restrict__ start, Particles P, TileJob job, FarMovers far) {
    QPIC_KERNEL_BEGIN(KID_COUNT + 1);
    QPIC_DYN_SMEM(double, acc);                                   // [4][TILE_CELLS]
    int * cnt = reinterpret_cast<int *>(acc + 4 * TILE_CELLS);    // [TILE_CELLS]
    int * lst = cnt + TILE_CELLS;                                 // [BIN_K][TILE_CELLS]
    const int sy = threadIdx.y, sz = threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y, tid = sy * blockDim.x + sz;
    const int j0 = ((job.tile_row0 + (int)blockIdx.y) % job.tile_rows) * TILE, k0 = blockIdx.x * TILE;
    const int ju = j0 - 2 + sy, ku = k0 - 2 + sz;  // unwrapped cell of this thread
    const int cell = wrap_near(ju, g.ny) * g.nz + wrap_near(ku, g.nz);
    const double * __restrict__ val = job.o[0].val;
    for (int i = tid; i < TILE_CELLS; i += nthreads) cnt[i] = 0;
    const int s0 = start[cell], s1 = start[cell + 1];
    const double ion = tile_prefetch_ion<1>(g, job, sy - 1, sz - 1, ju, ku);
    const bool is_target = sy >= 1 && sy <= TILE_ACC && sz >= 1 && sz <= TILE_ACC;
    // far arrivals of this (target) cell, listed by the predictor; almost always none
    const int nfar = (far.cnt && is_target) ? far.cnt[cell] : 0;
    if (far.clear && is_target && sy >= 2 && sz >= 2 && ju < g.ny && ku < g.nz) far.clear[cell] = 0;  // the node owner of the cell clears the other buffer
    // own particles, first four: everything a visit needs, one round trip
    int code[4];
    double pfy[4], pfz[4], pv[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const bool have = s0 + u < s1;
        code[u] = have ? P.off_m[s0 + u] : STAY_CODE;
        const double2 fm = have ? P.fyz_m[s0 + u] : make_double2(0.0, 0.0);
        pfy[u] = fm.x;
        pfz[u] = fm.y;
        pv[u] = have ? val[s0 + u] : 0.0;
    }
    __syncthreads();  // counters are zero

    // ---- phase A: movers of this source cell -> list of their target cell
    auto bin = [&](int s, int c) {
        if ((c & 63) == STAY_CODE || (c & 63) == FAR_CODE) return;
        const int ty = sy - 1 + (c & 7) - 2, tz = sz - 1 + ((c >> 3) & 7) - 2;
        if (ty >= 0 && ty < TILE_ACC && tz >= 0 && tz < TILE_ACC) {
            const int t = ty * TILE_ACC + tz;
            const int pos = atomicAdd(&cnt[t], 1);
            if (pos < BIN_K) lst[pos * TILE_CELLS + t] = (s << 2) | (((c >> 6) & 1) << 1) | ((c >> 7) & 1);
        }
    };
#pragma unroll
    for (int u = 0; u < 4; u++) {
        if (s0 + u < s1) bin(s0 + u, code[u]);
    }
    for (int s = s0 + 4; s < s1; s++) bin(s, P.off_m[s]);
    // far arrivals join the list of their target cell (this thread's own cell)
    for (int i = 0; i < nfar && i < FAR_K; i++) {
        const int pos = atomicAdd(&cnt[(sy - 1) * TILE_ACC + sz - 1], 1);
        if (pos < BIN_K) lst[pos * TILE_CELLS + (sy - 1) * TILE_ACC + sz - 1] = far.lst[(size_t)i * g.n2 + cell];
    }
    __syncthreads();

    // ---- phase B: ordered accumulation per target cell
    const int ty = sy - 1, tz = sz - 1;
    const int tcell = ty * TILE_ACC + tz;
    if (is_target) {
        double c00 = 0.0, c10 = 0.0, c01 = 0.0, c11 = 0.0;
        const double sign = job.o[0].sign;
        auto add = [&](double fy, double fz, double v) {
            c00 += v * (1 - fy) * (1 - fz);
            c10 += v * fy * (1 - fz);
            c01 += v * (1 - fy) * fz;
            c11 += v * fy * fz;
        };
        auto visit = [&](int s) {
            const double2 fm = P.fyz_m[s];
            add(fm.x, fm.y, sign * val[s]);
        };
        const int n = cnt[tcell];
        if (job.bin_cnt) job.bin_cnt[cell] = n;  // (halo cells: same value from the neighbour block)
        int * mine = lst + tcell;                // entry i at mine[i * TILE_CELLS]
        if (n <= BIN_K) {
            for (int i = 1; i < n; i++) {  // insertion sort: ascending slot = canonical order
                const int e = mine[i * TILE_CELLS];
                int q = i - 1;
                while (q >= 0 && mine[q * TILE_CELLS] > e) {
                    mine[(q + 1) * TILE_CELLS] = mine[q * TILE_CELLS];
                    q--;
                }
                mine[(q + 1) * TILE_CELLS] = e;
            }
            int i = 0;
            for (; i < n; i++) {  // movers from lower cells
                const int e = mine[i * TILE_CELLS];
                if ((e >> 2) >= s0) break;
                visit(e >> 2);
                if (job.bin_cnt) job.bin_lst[(size_t)i * g.n2 + cell] = e;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {  // stayers, from registers
                if (s0 + u < s1 && (code[u] & 63) == STAY_CODE) add(pfy[u], pfz[u], sign * pv[u]);
            }
            for (int s = s0 + 4; s < s1; s++) {
                if ((P.off_m[s] & 63) == STAY_CODE) visit(s);
            }
            for (; i < n; i++) {  // movers from higher cells
                const int e = mine[i * TILE_CELLS];
                visit(e >> 2);
                if (job.bin_cnt) job.bin_lst[(size_t)i * g.n2 + cell] = e;
            }
        } else {
            // crowded target (strong compression): scan the source window in order (stayers included).  The window does not see far
            // arrivals: a crowded cell that has some takes the rollback.
            if (nfar > 0) atomicExch(&ctl->fast_fail, 1);
            for (int dy_ = -1; dy_ <= 1; dy_++) {
                for (int dz_ = -1; dz_ <= 1; dz_++) {
                    const int src = mod_wrap(ju + dy_, g.ny) * g.nz + mod_wrap(ku + dz_, g.nz);
                    const int e1 = start[src + 1];
                    for (int s = start[src]; s < e1; s++) {
                        const int c = P.off_m[s];
                        if (dy_ + (c & 7) - 2 == 0 && dz_ + ((c >> 3) & 7) - 2 == 0) visit(s);
                    }
                }
            }
        }
        acc[0 * TILE_CELLS + tcell] = c00;
        acc[1 * TILE_CELLS + tcell] = c10;
        acc[2 * TILE_CELLS + tcell] = c01;
        acc[3 * TILE_CELLS + tcell] = c11;
    }
    QPIC_KERNEL_TAIL();
    __syncthreads();
    if (cancelled(ctl)) return;
    if (is_target) tile_store_nodes<1>(ctl, g, job, acc, ty, tz, ju, ku, ion);
}

// ------------------------------------------------------------------------------------------------ psi_middle source from the predictor
// EXPERIMENTAL (QPIC_PREDICT_DEPOSIT=1, single GPU): the predictor's lanes hold consecutive slots, i.e. the particles of a cell
// sit in neighbouring lanes, so the stayers of a cell can be summed where they are produced: k_particles_predict_dep does a
// warp-segmented sum of the four corner contributions (key = slice-start cell; a shuffle tree in lane order: deterministic) and
// the last lane of every run writes it -- to corn[c] when the cell's particles lie inside this warp, else to the warp's leading /
// trailing piece.  Movers are appended to the per-cell lists with integer atomics.  k_tile_assemble then needs one round trip per
// cell (slot range, mover count, corner sums) instead of the particle data: it chains the pieces, adds the (sorted) movers around
// them, publishes the sorted mover lists for k_tile_cell_deposit and forms the nodes.
// Measured (B200, in-graph timeline gpurun_out/r38_timeline_*.txt, config #2): k_tile_assemble 8.8 us against 15.6 us for
// k_tile_stay_deposit, but the predictor grows from 29.8 to 36.6 us (it is bandwidth- AND issue-bound: ~100 more instructions and 32
// more bytes per cell) and the jy/jz deposit from 23.6 to 26.8 us (the particle data it reads is no longer warmed in L2 by the
// deposit before it): 0.700 vs 0.696 ms/slice -- a wash, so it stays off.  Config #3 (1024^2, M = 10) gains: 4.51 vs 4.68 ms/slice.
struct StayOut {
    double * corn;   // [4][n2]: corner sums (c00, c10, c01, c11) of the stayers of cells whose run lies inside one warp
    double * lead;   // [nwarps][4]: run that continues from the previous warp (may also continue into the next one)
    double * trail;  // [nwarps][4]: run that starts in this warp and continues into the next one
    // FOLD_JX: the jx deposit at (y, z) (src/System_3d.cpp:511-527) rides along -- every particle is a "stayer" of its slice-start cell there, the
    // corrector's value n vx / (1 - vx) and the fractions are in the predictor's registers: same three arrays, sums of v (1-fy)(1-fz) ... v fy fz
    double *xcorn, *xlead, *xtrail;
    // FOLD 2 (last pass of a slice): rho and susceptibility(i) (:520-527) likewise, 8 sums: n / (1 - vx) and -n / (1 - vx) / gamma times the four weights
    double *rcorn, *rlead, *rtrail;
    int * cnt;       // movers per target cell (zero on entry)
    int * lst;       // [BIN_K][n2], unsorted
    double sign;     // -1: the psi source deposits -n (src/System_3d.cpp:349-353)
    int allow_far;   // half steps beyond one cell are listed like the other movers; 0: they raise Ctl::fast_fail
    int far_row_lo, far_row_n;  // slabs: legal target rows of a far mover, see FarMovers (0 rows: all)
};

#ifndef QPIC_PDEP_MINB
#define QPIC_PDEP_MINB 1
#endif

// SRED: how the per-cell sums leave the warp.  false: shuffle tree over the lanes of a run (as many steps as the longest run needs) -- the
// cheaper one for short runs (4 particles per cell: 551 vs 567 us/slice on config #2, 3514 vs 3629 on config #3); true: through shared memory,
// lane (k, r) adds value k over run r left to right -- independent of the run length, the cheaper one for dense cells (16 per cell: 91.7 vs
// 94.8 ms/slice on config #4).  Chosen by the host from the density (>= 8 particles per cell).
template <bool FIRST = false, int FOLD = 0, bool SRED = false>
__global__ void
#if QPIC_PDEP_MINB > 1
    __launch_bounds__(128, QPIC_PDEP_MINB)
#endif
    k_particles_predict_dep(Ctl * ctl, Geom g, Particles P, const double * __restrict__ by, const double * __restrict__ bz, int store_px,
                                        StayOut so, const double * __restrict__ a_sqr3, const double * __restrict__ psi) {
    QPIC_KERNEL_BEGIN(KID_k_particles_predict);
    constexpr bool FOLD_JX = FOLD >= 1;
    const int s_raw = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int live = live_count(ctl, P.np);  // (slab-decomposed sweeps: the live particles of this rank)
    if (s_raw - lane >= live) return;        // whole warp beyond the live range (uniform per warp)
    const bool have = s_raw < live;
    const int s = have ? s_raw : live - 1;  // the tail lanes of the last warp stay for the shuffles and repeat the last particle
    const int warp_base = s_raw - lane;
    const int c = P.cell_b[s];
    // the cells of the slots next to this warp: does its first / last run continue beyond it?
    const int key_before = warp_base > 0 ? P.cell_b[warp_base - 1] : -2;
    const int key_after = warp_base + 32 < live ? P.cell_b[warp_base + 32] : -2;
    const int jc = c / g.nz, kc = c - jc * g.nz;
    const double2 fyz = P.fyz_b[s];
    const double fy = fyz.x, fz = fyz.y;
    const double2 pyz = ld_stream2(P.pyz_s + s);
    const double py = pyz.x, pz = pyz.y;
    const PredictIn in = predict_inputs<FIRST>(ctl, g, P, s, jc, kc, fy, fz, py, pz, a_sqr3, psi);
    const double q = in.q, gamma = in.gamma, a_particle = in.a_particle;
    const double t1y = in.t1y, t1z = in.t1z, dpy = in.dpy, dpz = in.dpz;
    const double y0 = slot_position(g.exact_dy, jc, fy, g.dy, P.y_s, s), z0 = slot_position(g.exact_dz, kc, fz, g.dz, P.z_s, s);
    const double nw = (QPIC_L2_KEEP & 4) ? ld_keep1(P.nw_s + s) : P.nw_s[s];
    int js, ks;
    double fys, fzs;
    stagger(jc, fy, g.ny, js, fys);
    stagger(kc, fz, g.nz, ks, fzs);
    const double by_particle = gather4(by, g.nz, jc, wrap_inc(jc, g.ny), ks, wrap_inc(ks, g.nz), fy, fzs);
    const double bz_particle = gather4(bz, g.nz, js, wrap_inc(js, g.ny), kc, wrap_inc(kc, g.nz), fys, fz);
    if (cancelled(ctl)) return;  // uniform over the grid
    QPIC_KERNEL_TAIL();
    // ---- exactly k_particles_predict<true> from here ...
    const double rq = 1.0 / q;  // (see k_particles_gamma)
    double pyn = py - t1y;
    pyn += g.dx * gamma * dpy * rq;
    pyn -= g.dx * bz_particle;
    double pzn = pz - t1z;
    pzn += g.dx * gamma * dpz * rq;
    pzn += g.dx * by_particle;
    const double p_squared = 0.25 * (py + pyn) * (py + pyn) + 0.25 * (pz + pzn) * (pz + pzn);
    const double gamma_c = 0.5 * (1 + p_squared + a_particle + q * q) * rq;
    const double px_c = 0.5 * (1 + p_squared + a_particle - q * q) * rq;
    const double ym = wrap_fast(y0 + 0.5 * g.dx * pyn * rq, g.ly);
    const double zm = wrap_fast(z0 + 0.5 * g.dx * pzn * rq, g.lz);
    int jm, km;
    double f1, f2;
    locate_fast(ym, g.dy, g.inv_dy, g.exact_dy, g.ny, jm, f1);
    locate_fast(zm, g.dz, g.inv_dz, g.exact_dz, g.nz, km, f2);
    const int oy = cell_offset(jm, jc, g.ny), oz = cell_offset(km, kc, g.nz);
    const bool bad_position = !(f1 >= 0.0) || !(f2 >= 0.0);  // NaN positions of a blown-up run
    // a half step beyond one cell: listed under its target cell like every mover of this kernel, but marked (offset code FAR, upper half
    // of the counter) because the in-order window scan that serves crowded cells would not find it
    const bool beyond = oy < -1 || oy > 1 || oz < -1 || oz > 1;
    const bool bad = bad_position || (beyond && !(so.allow_far && far_row_ok(so.far_row_lo, so.far_row_n, jm, g.ny)));
    const bool farm = !bad && beyond;
    const int code = bad ? pack_offset(0, 0, 0.0, 0.0, kc)
                         : (farm ? (FAR_CODE | ((f1 >= 0.5) ? 64 : 0) | ((f2 >= 0.5) ? 128 : 0) | ((kc & 3) << 8)) : pack_offset(oy, oz, f1, f2, kc));
    if (have) {
        predict_store_invariants<FIRST>(P, s, in);
        st_keep(P.pyzn + s, make_double2(pyn, pzn));
        P.gam[s] = gamma_c;
        if (store_px) P.pxs[s] = px_c;
        if (!FOLD_JX) store_currents<false>(P, s, nw, gamma_c, px_c);  // (FOLD_JX: no jx tile deposit reads the value)
        st_keep(P.fyz_m + s, make_double2(f1, f2));
        st_keep(P.off_m + s, code);
        if (bad) atomicExch(&ctl->fast_fail, 1);
    }
    // ---- ... and the psi_middle source: stayers summed per cell, movers listed
    const bool stay = have && !bad && oy == 0 && oz == 0;
    if (have && !bad && !stay) {
        const int target = jm * g.nz + km;
        const int pos = atomicAdd(&so.cnt[target], farm ? 0x10001 : 1) & 0xffff;
        if (pos < BIN_K) so.lst[(size_t)pos * g.n2 + target] = (s << 2) | (((code >> 6) & 1) << 1) | ((code >> 7) & 1);
        else if (farm) atomicExch(&ctl->fast_fail, 1);
    }
    const double v = stay ? so.sign * nw : 0.0;
    constexpr int NV = FOLD == 2 ? 16 : (FOLD == 1 ? 8 : 4);
    double cs[NV];
    cs[0] = v * (1 - f1) * (1 - f2), cs[1] = v * f1 * (1 - f2), cs[2] = v * (1 - f1) * f2, cs[3] = v * f1 * f2;
    if (!stay) cs[0] = cs[1] = cs[2] = cs[3] = 0.0;  // (NaN positions of a blown-up run)
    if (FOLD_JX) {
        // store_currents' value, k_tile_deposit's products (at the slice-start fractions)
        const double vx = px_c / gamma_c;
        const double vj = have ? nw * vx / (1 - vx) : 0.0;
        cs[4] = vj * (1 - fy) * (1 - fz), cs[5] = vj * fy * (1 - fz), cs[6] = vj * (1 - fy) * fz, cs[7] = vj * fy * fz;
        if (FOLD == 2) {
            // k_tile_deposit<2, ., true>'s values
            const double r = have ? nw / (1 - vx) : 0.0;
            const double ch = -r / gamma_c;
            cs[8] = r * (1 - fy) * (1 - fz), cs[9] = r * fy * (1 - fz), cs[10] = r * (1 - fy) * fz, cs[11] = r * fy * fz;
            cs[12] = ch * (1 - fy) * (1 - fz), cs[13] = ch * fy * (1 - fz), cs[14] = ch * (1 - fy) * fz, cs[15] = ch * fy * fz;
        }
    }
    const int key = have ? c : -3 - lane;
    if constexpr (SRED) {
    // Segmented sums through shared memory: every lane parks its NV contributions in the warp's tile; lane (k, r) then adds value k over run r
    // in slot order (runs = maximal groups of lanes with the same cell; 32 / NV runs per round).  Same association order as the tile kernels
    // (a plain left-to-right sum per cell), ~half the instructions of the shuffle tree, and independent of the run length.
    {
        constexpr int WARPS = 4;  // QPIC_PBLOCK / 32
        __shared__ double red[WARPS][NV][33];
        __shared__ int run_start[WARPS][34];
        const int wib = (threadIdx.x >> 5) & (WARPS - 1);
#pragma unroll
        for (int i = 0; i < NV; i++) red[wib][i][lane] = cs[i];
        const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool starts = lane == 0 || key_prev != key;
        const unsigned smask = __ballot_sync(0xffffffffu, starts);
        const int nruns = __popc(smask);
        if (starts) run_start[wib][__popc(smask & ((1u << lane) - 1u))] = lane;
        if (lane == 0) run_start[wib][nruns] = 32;
        __syncwarp();
        constexpr int RPR = 32 / NV;  // runs per round
        const int k = lane % NV, rr = lane / NV;
        const int w = warp_base >> 5;
        const int key_first = __shfl_sync(0xffffffffu, key, 0), key_last = __shfl_sync(0xffffffffu, key, 31);
        for (int base = 0; base < nruns; base += RPR) {
            const int r = base + rr;
            const bool mine = r < nruns;
            const int st = mine ? run_start[wib][r] : 0, en = mine ? run_start[wib][r + 1] : 0;
            const int kr = __shfl_sync(0xffffffffu, key, st);  // the run's cell (negative: lanes beyond the live range)
            if (!mine || kr < 0) continue;
            double sum = red[wib][k][st];
            for (int q = st + 1; q < en; q++) sum += red[wib][k][q];
            const bool from_before = st == 0 && kr == key_first && kr == key_before;
            const bool into_after = en == 32 && kr == key_last && kr == key_after;
            double * dst;
            if (from_before) dst = k < 4 ? so.lead + 4 * w + k : (k < 8 ? so.xlead + 4 * w + (k - 4) : so.rlead + 8 * w + (k - 8));
            else if (into_after) dst = k < 4 ? so.trail + 4 * w + k : (k < 8 ? so.xtrail + 4 * w + (k - 4) : so.rtrail + 8 * w + (k - 8));
            else dst = k < 4 ? so.corn + k * (size_t)g.n2 + kr : (k < 8 ? so.xcorn + (k - 4) * (size_t)g.n2 + kr : so.rcorn + (k - 8) * (size_t)g.n2 + kr);
            *dst = sum;
        }
    }
    } else {
    // keys are sorted, so once no lane finds its key d lanes below, none will at a larger distance: runs of 4-8 particles need 3 steps
    for (int d = 1; d < 32; d <<= 1) {
        const int k2 = __shfl_up_sync(0xffffffffu, key, d);
        if (!__any_sync(0xffffffffu, lane >= d && k2 == key)) break;
        const bool take = lane >= d && k2 == key;
#pragma unroll
        for (int i = 0; i < NV; i++) {
            const double t = __shfl_up_sync(0xffffffffu, cs[i], d);
            if (take) cs[i] += t;
        }
    }
    const int key_next = __shfl_down_sync(0xffffffffu, key, 1);
    const int key_first = __shfl_sync(0xffffffffu, key, 0);
    if (have && (lane == 31 || key_next != key)) {
        const bool from_before = key == key_first && key == key_before;
        const bool into_after = lane == 31 && key == key_after;
        const int w = warp_base >> 5;
        if (from_before) {
#pragma unroll
            for (int i = 0; i < 4; i++) so.lead[4 * w + i] = cs[i];
            if (FOLD_JX) {
#pragma unroll
                for (int i = 0; i < 4; i++) so.xlead[4 * w + i] = cs[4 + i];
            }
            if (FOLD == 2) {
#pragma unroll
                for (int i = 0; i < 8; i++) so.rlead[8 * w + i] = cs[8 + i];
            }
        } else if (into_after) {
#pragma unroll
            for (int i = 0; i < 4; i++) so.trail[4 * w + i] = cs[i];
            if (FOLD_JX) {
#pragma unroll
                for (int i = 0; i < 4; i++) so.xtrail[4 * w + i] = cs[4 + i];
            }
            if (FOLD == 2) {
#pragma unroll
                for (int i = 0; i < 8; i++) so.rtrail[8 * w + i] = cs[8 + i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (QPIC_L2_KEEP & 8) st_keep1(so.corn + i * (size_t)g.n2 + c, cs[i]);
                else so.corn[i * (size_t)g.n2 + c] = cs[i];
            }
            if (FOLD_JX) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (QPIC_L2_KEEP & 8) st_keep1(so.xcorn + i * (size_t)g.n2 + c, cs[4 + i]);
                    else so.xcorn[i * (size_t)g.n2 + c] = cs[4 + i];
                }
            }
            if (FOLD == 2) {
#pragma unroll
                for (int i = 0; i < 8; i++) so.rcorn[i * (size_t)g.n2 + c] = cs[8 + i];
            }
        }
    }
    }
}

// nodes of the psi_middle source from what k_particles_predict_dep left; block = TILE_ACC x TILE_ACC target cells as in k_tile_deposit.
// cnt_next: the mover counters of the next iteration's predictor (double buffered) are cleared by the cells' owner threads.
// FOLD_JX: second output = jx at (y, z) from the predictor's jx sums (no movers there: a cell's sum is its corner sum or its chain of pieces)
// Blocks per SM the compiler must leave room for: 4 = 48 registers.  Measured on config #2 (us/slice, gpurun_out/r75_ab.log): unspecified (58
// registers, 3 blocks) 561, 4 -> 550, 5 / 6 (32 registers, spills) 588 / 584; an explicit 1 lets the compiler take 104 registers: 597.
#ifndef QPIC_ASM_MINB
#define QPIC_ASM_MINB 4
#endif
template <bool FOLD_JX = false>
__global__ void __launch_bounds__(TILE_CELLS, QPIC_ASM_MINB) k_tile_assemble(Ctl * ctl, Geom g, const int * __restrict__ start, Particles P, TileJob job, StayOut so,
                                                              int * __restrict__ cnt_next) {
    QPIC_KERNEL_BEGIN(KID_COUNT + 1);
    __shared__ double acc[(FOLD_JX ? 8 : 4) * TILE_CELLS];
    const int ay = threadIdx.y, az = threadIdx.x;
    const int tcell = ay * TILE_ACC + az;
    const int jt = ((job.tile_row0 + (int)blockIdx.y) % job.tile_rows) * TILE - 1 + ay, kt = blockIdx.x * TILE - 1 + az;  // unwrapped target cell
    const int cell = wrap_near(jt, g.ny) * g.nz + wrap_near(kt, g.nz);
    const bool owner = ay >= 1 && az >= 1 && jt < g.ny && kt < g.nz;
    const int s0 = start[cell], s1 = start[cell + 1];
    const int nraw = so.cnt[cell];
    const int n = nraw & 0xffff, nfar = nraw >> 16;  // movers that arrived in this cell / how many of them from beyond the +-1 window
    const double ion = tile_prefetch_ion<1>(g, job, ay, az, jt, kt);
    const double k00 = so.corn[0 * (size_t)g.n2 + cell], k10 = so.corn[1 * (size_t)g.n2 + cell], k01 = so.corn[2 * (size_t)g.n2 + cell],
                 k11 = so.corn[3 * (size_t)g.n2 + cell];
    const double * __restrict__ val = job.o[0].val;
    double c00 = 0.0, c10 = 0.0, c01 = 0.0, c11 = 0.0;
    double x00 = 0.0, x10 = 0.0, x01 = 0.0, x11 = 0.0;
    if (FOLD_JX && s1 > s0) {
        const int w0 = s0 >> 5, w1 = (s1 - 1) >> 5;
        if (w0 == w1) {
            x00 = so.xcorn[0 * (size_t)g.n2 + cell], x10 = so.xcorn[1 * (size_t)g.n2 + cell], x01 = so.xcorn[2 * (size_t)g.n2 + cell], x11 = so.xcorn[3 * (size_t)g.n2 + cell];
        } else {
            x00 = so.xtrail[4 * w0 + 0], x10 = so.xtrail[4 * w0 + 1], x01 = so.xtrail[4 * w0 + 2], x11 = so.xtrail[4 * w0 + 3];
            for (int w = w0 + 1; w <= w1; w++) x00 += so.xlead[4 * w + 0], x10 += so.xlead[4 * w + 1], x01 += so.xlead[4 * w + 2], x11 += so.xlead[4 * w + 3];
        }
    }
    auto visit = [&](int s) {
        const double2 fm = P.fyz_m[s];
        const double fy = fm.x, fz = fm.y, v = so.sign * val[s];
        c00 += v * (1 - fy) * (1 - fz);
        c10 += v * fy * (1 - fz);
        c01 += v * (1 - fy) * fz;
        c11 += v * fy * fz;
    };
    auto stayers = [&]() {
        if (s1 <= s0) return;
        const int w0 = s0 >> 5, w1 = (s1 - 1) >> 5;
        if (w0 == w1) {
            c00 += k00, c10 += k10, c01 += k01, c11 += k11;
        } else {
            c00 += so.trail[4 * w0 + 0], c10 += so.trail[4 * w0 + 1], c01 += so.trail[4 * w0 + 2], c11 += so.trail[4 * w0 + 3];
            for (int w = w0 + 1; w <= w1; w++) c00 += so.lead[4 * w + 0], c10 += so.lead[4 * w + 1], c01 += so.lead[4 * w + 2], c11 += so.lead[4 * w + 3];
        }
    };
    if (n <= BIN_K) {
        int e[BIN_K];
        for (int i = 0; i < n; i++) e[i] = so.lst[(size_t)i * g.n2 + cell];
        for (int i = 1; i < n; i++) {  // insertion sort: ascending slot
            const int x = e[i];
            int q = i - 1;
            while (q >= 0 && e[q] > x) {
                e[q + 1] = e[q];
                q--;
            }
            e[q + 1] = x;
        }
        int i = 0;
        for (; i < n && (e[i] >> 2) < s0; i++) visit(e[i] >> 2);
        stayers();
        for (; i < n; i++) visit(e[i] >> 2);
        if (owner && job.bin_lst) {
            for (int q = 0; q < n; q++) job.bin_lst[(size_t)q * g.n2 + cell] = e[q];
        }
    } else {
        // crowded target (strong compression): scan the source window in order, stayers included (far arrivals are not in the window:
        // a crowded cell that has some takes the rollback)
        if (nfar > 0) atomicExch(&ctl->fast_fail, 1);
        for (int dy_ = -1; dy_ <= 1; dy_++) {
            for (int dz_ = -1; dz_ <= 1; dz_++) {
                const int src = mod_wrap(jt + dy_, g.ny) * g.nz + mod_wrap(kt + dz_, g.nz);
                const int e1 = start[src + 1];
                for (int s = start[src]; s < e1; s++) {
                    const int cd = P.off_m[s];
                    if (dy_ + (cd & 7) - 2 == 0 && dz_ + ((cd >> 3) & 7) - 2 == 0) visit(s);
                }
            }
        }
    }
    if (cancelled(ctl)) return;
    if (owner) {
        if (job.bin_cnt) job.bin_cnt[cell] = n;
        if (cnt_next) cnt_next[cell] = 0;
    }
    QPIC_KERNEL_TAIL();
    acc[0 * TILE_CELLS + tcell] = c00;
    acc[1 * TILE_CELLS + tcell] = c10;
    acc[2 * TILE_CELLS + tcell] = c01;
    acc[3 * TILE_CELLS + tcell] = c11;
    if (FOLD_JX) {
        acc[4 * TILE_CELLS + tcell] = x00;
        acc[5 * TILE_CELLS + tcell] = x10;
        acc[6 * TILE_CELLS + tcell] = x01;
        acc[7 * TILE_CELLS + tcell] = x11;
    }
    __syncthreads();
    tile_store_nodes<FOLD_JX ? 2 : 1>(ctl, g, job, acc, ay, az, jt, kt, ion);
}

// rho and susceptibility(i) of the slice from the sums the LAST predictor pass left (k_particles_predict_dep<., 2>): no particle is read.
// Runs like k_tile_deposit<2, true, true> (FINAL_PASS) -- except when the |B| stop condition ended the slice before the last pass ran: the sums
// do not exist then and the tile kernel, which is launched right after this one with TileJob::rc_folded set, does the work instead.
__global__ void __launch_bounds__(TILE_CELLS) k_tile_fold_rho_chi(const Ctl * ctl, Geom g, const int * __restrict__ start, TileJob job, StayOut so, int last_it) {
    QPIC_KERNEL_BEGIN(KID_COUNT + 11);
    __shared__ double acc[8 * TILE_CELLS];
    const int ay = threadIdx.y, az = threadIdx.x;
    const int tcell = ay * TILE_ACC + az;
    const int jt = ((job.tile_row0 + (int)blockIdx.y) % job.tile_rows) * TILE - 1 + ay, kt = blockIdx.x * TILE - 1 + az;  // unwrapped target cell
    const int cell = wrap_near(jt, g.ny) * g.nz + wrap_near(kt, g.nz);
    const int s0 = start[cell], s1 = start[cell + 1];
    double r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = 0.0;
    if (s1 > s0) {
        const int w0 = s0 >> 5, w1 = (s1 - 1) >> 5;
        if (w0 == w1) {
#pragma unroll
            for (int i = 0; i < 8; i++) r[i] = so.rcorn[i * (size_t)g.n2 + cell];
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) r[i] = so.rtrail[8 * w0 + i];
            for (int w = w0 + 1; w <= w1; w++) {
#pragma unroll
                for (int i = 0; i < 8; i++) r[i] += so.rlead[8 * w + i];
            }
        }
    }
    const int tag = ctl->stop_tag;
    if (ctl->fast_fail) return;
    if (tag != 0 && ((tag - 1) / STOP_TAG_STRIDE != ctl->slice || (tag - 1) % STOP_TAG_STRIDE != last_it)) return;
    QPIC_KERNEL_TAIL();
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i * TILE_CELLS + tcell] = r[i];
    __syncthreads();
    tile_store_nodes<2>(ctl, g, job, acc, ay, az, jt, kt, 0.0);
}

// ------------------------------------------------------------------------------------------------ jy / jz, one visit per particle
// Successor of round 1's three-list kernel (k_tile_list_deposit, removed).  There a thread owns a STAGGERED target cell of both outputs and
// walks three lists -- its own cell's and those of the cells above / to the right, filtering by the half-cell bits -- so every
// list entry is scanned three times and a particle's 1 / (1 + psi_middle) is formed up to three times.  Here a thread owns the
// UNSTAGGERED half-step cell (j, k), walks only that cell's list, and every particle is visited exactly once: its value
// n p_next / (1 + psi_middle@particle) is formed once for both outputs and scattered to the thread's private NODE sums --
// jy nodes (j-1 .. j+1) x (k, k+1) and jz nodes (j, j+1) x (k-1 .. k+1), 12 registers -- in list (= ascending slot) order.
// After a barrier node (J, K) adds the six sums the cells around it hold for it, in a fixed order.  Deterministic; per
// particle the products are the reference's (value * wy) * wz (src/array_utils.cpp:75-82 on the staggered grids of
// src/System_3d.cpp:190-193).  Measured on config #2: own-list-only work of the old kernel was 11 us of its 29.5 us.
// Block = (LIST_TY + 2) x (LIST_TZ + 2) = 16 x 18 = 288 threads (9 full warps): thread (ay, az) <-> cell (j0 - 1 + ay, k0 - 1 + az),
// nodes of the block = cells with 1 <= ay <= LIST_TY, 1 <= az <= LIST_TZ.
// Two tile widths (template argument TZ): 14 x 16 nodes (288 threads, two blocks per SM) and, for dense cells (>= 8 particles per cell), 14 x 32
// nodes (544 threads, one block per SM): 21 % instead of 29 % of the visits are halo cells' -- config #4 on one GPU 91.7 -> 89.7 ms/slice; at 4
// particles per cell the wide tile gains nothing (550.5 vs 549.8 us/slice on config #2, 3517 vs 3484 on config #3).
constexpr int LIST_TY = 14, LIST_BY = LIST_TY + 2;

// Measured on config #2 (gpurun_out, third session; whole-slice times): 3 blocks / SM (72 registers, spills) 0.742 ms, with the
// two-deep pipeline 0.781 (more spills); 2 blocks / SM (96 registers, no spills) 0.731, with the pipeline 0.722.  The three-list
// kernel: 0.762.  With the stayers-direct lists (0.704): stayers four at a time 0.744, epilogue operands prefetched at kernel start
// 0.718 (spills), both 0.756 -- more registers per thread lose here even without spills.
#ifndef QPIC_CELL_MINB
#define QPIC_CELL_MINB 2
#endif
#ifndef QPIC_CELL_PIPE
#define QPIC_CELL_PIPE 1
#endif
#ifndef QPIC_CELL_UNROLL
#define QPIC_CELL_UNROLL 2
#endif
// L1 prefetch of the B-source epilogue's operands at kernel start: 551 -> 549.5 us/slice on config #2, 3514 -> 3482 on config #3
#ifndef QPIC_CELL_EPI_PREFETCH
#define QPIC_CELL_EPI_PREFETCH 1
#endif
template <int LIST_TZ>
__global__ void __launch_bounds__(LIST_BY * (LIST_TZ + 2), LIST_TZ == 16 ? QPIC_CELL_MINB : 1)
    k_tile_cell_deposit(const Ctl * ctl, Geom g, const int * __restrict__ start, Particles P, TileJob job) {
    constexpr int LIST_BZ = LIST_TZ + 2, LIST_THREADS = LIST_BY * LIST_BZ;
    QPIC_KERNEL_BEGIN(KID_COUNT + 5);
    QPIC_DYN_SMEM(double, acc);  // [12][LIST_THREADS]
    const int ay = threadIdx.y, az = threadIdx.x;
    const int tid = ay * LIST_BZ + az;
    // node rows [cell_row0, cell_row0 + cell_rows) are launched: the whole grid on a single GPU, the slab and one tile row on
    // either side in a slab-decomposed sweep (rows wrap; cell_row0 + cell_rows + LIST_TY < 2 ny)
    const int r0 = blockIdx.y * LIST_TY, k0 = blockIdx.x * LIST_TZ;
    const int jt = job.cell_row0 + r0 - 1 + ay, kt = k0 - 1 + az;  // unwrapped cell of this thread
    const int jw = wrap_near(jt, g.ny), kw = wrap_near(kt, g.nz);
    const int cell = jw * g.nz + kw;
    // movers exist only where the psi_middle source deposit of this rank published them
    int lr = jw - job.list_row0;
    lr += (lr < 0) ? g.ny : 0;
    const int n = lr < job.list_rows ? job.bin_cnt[cell] : 0;
    const double2 * __restrict__ pyzn = P.pyzn;
    const double2 * __restrict__ fyzm = P.fyz_m;
#if QPIC_CELL_EPI_PREFETCH && !defined(QPIC_EMU)
    // the epilogue's operands (node (jw, kw) of this thread) into L1 now: no registers held, and the loads after the barrier hit
    if (job.b_src && ay >= 1 && ay <= LIST_TY && az >= 1 && az <= LIST_TZ) {
        const int tn = jw * g.nz + kw;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(job.b_jz + tn));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(job.b_jy + tn));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(job.b_jx + tn));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(job.b_by + tn));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(job.b_bz + tn));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(job.b_jx + wrap_inc(jw, g.ny) * g.nz + kw));
    }
#endif
    // psi_middle at the four nodes of this cell: every particle of the list gathers from them (src/System_3d.cpp:480-482)
    const int jw1 = wrap_inc(jw, g.ny), kw1 = wrap_inc(kw, g.nz);
    const double p00 = job.psi_mid[jw * g.nz + kw], p10 = job.psi_mid[jw1 * g.nz + kw], p01 = job.psi_mid[jw * g.nz + kw1],
                 p11 = job.psi_mid[jw1 * g.nz + kw1];
    // jy node sums: ym* = row j-1, y0* = row j, yp* = row j+1; second index: column k / k+1.  jz: z*m = column k-1, z*0 = k, z*p = k+1.
    double ym0 = 0.0, ym1 = 0.0, y00 = 0.0, y01 = 0.0, yp0 = 0.0, yp1 = 0.0;
    double z0m = 0.0, z1m = 0.0, z00 = 0.0, z10 = 0.0, z0p = 0.0, z1p = 0.0;
    auto visit5 = [&](double fy, double fz, double nw, double vy0, double vz0, int hy, int hz) {
        // gather_node's expression on the preloaded corners
        const double rw = 1.0 / (1 + (p00 * (1 - fy) * (1 - fz) + p10 * fy * (1 - fz) + p01 * (1 - fy) * fz + p11 * fy * fz));
        const double vy = nw * vy0 * rw, vz = nw * vz0 * rw;  // one reciprocal instead of two quotients (within an ulp)
        {
            // jy: staggered by dy/2 in y -- the lower node row is j (upper half of the cell) or j-1 (lower half)
            const double wy = hy ? fy - 0.5 : fy + 0.5;
            const double c00 = vy * (1 - wy) * (1 - fz), c10 = vy * wy * (1 - fz), c01 = vy * (1 - wy) * fz, c11 = vy * wy * fz;
            ym0 += hy ? 0.0 : c00;
            ym1 += hy ? 0.0 : c01;
            y00 += hy ? c00 : c10;
            y01 += hy ? c01 : c11;
            yp0 += hy ? c10 : 0.0;
            yp1 += hy ? c11 : 0.0;
        }
        {
            const double wz = hz ? fz - 0.5 : fz + 0.5;
            const double c00 = vz * (1 - fy) * (1 - wz), c10 = vz * fy * (1 - wz), c01 = vz * (1 - fy) * wz, c11 = vz * fy * wz;
            z0m += hz ? 0.0 : c00;
            z1m += hz ? 0.0 : c10;
            z00 += hz ? c00 : c01;
            z10 += hz ? c10 : c11;
            z0p += hz ? c01 : 0.0;
            z1p += hz ? c11 : 0.0;
        }
    };
    auto visit = [&](int s, int hy, int hz) {
        const double2 fm = fyzm[s], pn = pyzn[s];
        visit5(fm.x, fm.y, job.nw[s], pn.x, pn.y, hy, hz);
    };
    if (n <= BIN_K) {
        // the list holds the MOVERS that ended their half step in this cell (k_tile_stay_deposit), sorted by slot; the cell's own
        // stayers are the particles of its slot range whose offset code is (0, 0).  Ascending slot order = movers from lower
        // cells, stayers, movers from higher cells.  Stayers are the bulk: their data is loaded two particles at a time.
        const int s0 = start[cell], s1 = start[cell + 1];
        int i = 0;
        int e = n > 0 ? job.bin_lst[cell] : 0;
        for (; i < n && (e >> 2) < s0; ) {
            const int ce = e;
            i++;
            if (i < n) e = job.bin_lst[(size_t)i * g.n2 + cell];
            visit(ce >> 2, (ce >> 1) & 1, ce & 1);
        }
        for (int sb = s0; sb < s1; sb += QPIC_CELL_UNROLL) {
            // QPIC_CELL_UNROLL particles per round trip (out-of-range ones repeat the first and are skipped)
            int cd[QPIC_CELL_UNROLL];
            double2 fm[QPIC_CELL_UNROLL], pn[QPIC_CELL_UNROLL];
            double nwv[QPIC_CELL_UNROLL];
#pragma unroll
            for (int u = 0; u < QPIC_CELL_UNROLL; u++) {
                const int sU = sb + u < s1 ? sb + u : sb;
                cd[u] = ld_release(P.off_m + sU);
                fm[u] = ld_release(fyzm + sU);
                pn[u] = ld_release(pyzn + sU);
                nwv[u] = job.nw[sU];
            }
#pragma unroll
            for (int u = 0; u < QPIC_CELL_UNROLL; u++) {
                if (sb + u < s1 && (cd[u] & 63) == STAY_CODE) visit5(fm[u].x, fm[u].y, nwv[u], pn[u].x, pn[u].y, (cd[u] >> 6) & 1, (cd[u] >> 7) & 1);
            }
        }
        for (; i < n; ) {
            const int ce = e;
            i++;
            if (i < n) e = job.bin_lst[(size_t)i * g.n2 + cell];
            visit(ce >> 2, (ce >> 1) & 1, ce & 1);
        }
    } else {
        // a crowded cell was not listed (strong compression): scan the source window in order
        for (int dy_ = -1; dy_ <= 1; dy_++) {
            for (int dz_ = -1; dz_ <= 1; dz_++) {
                const int src = wrap_near(jt + dy_, g.ny) * g.nz + wrap_near(kt + dz_, g.nz);
                const int e1 = start[src + 1];
                for (int s = start[src]; s < e1; s++) {
                    const int code = P.off_m[s];
                    if (dy_ + (code & 7) - 2 != 0 || dz_ + ((code >> 3) & 7) - 2 != 0) continue;
                    visit(s, (code >> 6) & 1, (code >> 7) & 1);
                }
            }
        }
    }
    if (cancelled(ctl)) return;
    QPIC_KERNEL_TAIL();
    acc[0 * LIST_THREADS + tid] = ym0;
    acc[1 * LIST_THREADS + tid] = ym1;
    acc[2 * LIST_THREADS + tid] = y00;
    acc[3 * LIST_THREADS + tid] = y01;
    acc[4 * LIST_THREADS + tid] = yp0;
    acc[5 * LIST_THREADS + tid] = yp1;
    acc[6 * LIST_THREADS + tid] = z0m;
    acc[7 * LIST_THREADS + tid] = z1m;
    acc[8 * LIST_THREADS + tid] = z00;
    acc[9 * LIST_THREADS + tid] = z10;
    acc[10 * LIST_THREADS + tid] = z0p;
    acc[11 * LIST_THREADS + tid] = z1p;
    __syncthreads();
    if (ay < 1 || ay > LIST_TY || az < 1 || az > LIST_TZ || r0 + ay - 1 >= job.cell_rows || kt >= g.nz) return;
    // node (jt, kt).  jy: what the cells of rows jt+1 / jt / jt-1 hold for their row j-1 / j / j+1, columns kt (their k) and kt-1
    // (their k+1); jz likewise with rows and columns exchanged.  Fixed order.
    const int t = jw * g.nz + kt;
    const int up = tid + LIST_BZ, dn = tid - LIST_BZ;
    double jyn = 0.0, jzn = 0.0;
    jyn += acc[2 * LIST_THREADS + tid];
    jyn += acc[3 * LIST_THREADS + tid - 1];
    jyn += acc[4 * LIST_THREADS + dn];
    jyn += acc[5 * LIST_THREADS + dn - 1];
    jyn += acc[0 * LIST_THREADS + up];
    jyn += acc[1 * LIST_THREADS + up - 1];
    jzn += acc[8 * LIST_THREADS + tid];
    jzn += acc[9 * LIST_THREADS + dn];
    jzn += acc[10 * LIST_THREADS + tid - 1];
    jzn += acc[11 * LIST_THREADS + dn - 1];
    jzn += acc[6 * LIST_THREADS + tid + 1];
    jzn += acc[7 * LIST_THREADS + dn + 1];
    job.o[0].out[t] = jyn;
    job.o[1].out[t] = jzn;
    if (job.b_src) {
        // same expressions as k_p2_rows_b_forward
        const int jp = wrap_inc(jw, g.ny), kp = wrap_inc(kt, g.nz);
        const int jq = (jw == g.ny - 1) ? 0 : jw;
        const double djz_dxi = (job.b_jz[t] - jzn) * g.inv_dx;
        const double djy_dxi = (job.b_jy[t] - jyn) * g.inv_dx;
        const double jxt = job.b_jx[t];
        job.b_src[t] = make_double2(-job.b_D * job.b_by[t] - (job.b_jx[jw * g.nz + kp] - jxt) * g.inv_dz + djz_dxi,
                                    -job.b_D * job.b_bz[jq * g.nz + kt] + (job.b_jx[jp * g.nz + kt] - jxt) * g.inv_dy - djy_dxi);
    }
}

}  // namespace qpic
