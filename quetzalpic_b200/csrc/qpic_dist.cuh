// Slab-decomposed multi-GPU sweep (BASELINE config #4: 4096 x 4096 transverse grid on 2/4/8 B200).
//
// The transverse grid is cut into G slabs of ny/G rows; rank r owns rows [row0, row0 + ny_loc) and the plasma particles
// whose slice-start cell lies in them.  Every rank keeps full-size (ny, nz) arrays with GLOBAL indices (own rows + two halo
// rows on either side are meaningful), so all single-GPU kernels run unchanged on the rows / tile rows they are pointed at.
// What crosses GPUs goes over NVLink as plain peer loads / stores into an "exchange region" that every rank allocates with
// the same layout and maps into its peers (CUDA IPC between processes, raw pointers inside one process):
//   * transform transposes  -- the forward row kernels STORE each spectrum element straight into the column buffer of the
//                              rank that owns its column (coalesced along k); the backward row kernels LOAD their rows
//                              from the peers' column buffers (coalesced along k).  No pack / unpack kernels, no staging.
//   * halo reduce           -- partial deposit sums of the two rows beyond either slab edge go to the neighbour, which adds
//                              them in a fixed order (own + from below + from above): deterministic.
//   * halo fill             -- the two edge rows of psi, psi_middle, By, Bz, jx are written into the neighbours' arrays.
//   * particle migration    -- particles whose new cell belongs to a neighbour are appended to its inbox.
//   * k_xchg_barrier        -- flag exchange with system-scope release / acquire; one per exchange step.  All ranks issue the
//                              same sequence of barriers on ONE stream, so a device-side epoch counter pairs them up and the
//                              whole slice stays capturable in a CUDA graph.  Spins are bounded (error flag, never a hang).
#pragma once

#include "qpic_kernels.cuh"
#ifdef QPIC_EMU
#include <time.h>
#endif

namespace qpic {

constexpr int DIST_MAX = 8;
// halo rows on either side of a slab: 2 suffice for the gathers and the +-1 cell movers; 4 let half steps of up to three cells cross a
// slab edge on the sort-free path (far movers, qpic_kernels.cuh) -- the halo reduce carries their partial sums to the owner
constexpr int DIST_HALO = 4;
// reduce buffers, one per deposit target so that a buffer is reused only a whole iteration later
enum DistReduceSlot { RB_CHARGE = 0, RB_JX, RB_JY, RB_JZ, RB_RHO, RB_CHI, RB_COUNT };
// fields that live inside the exchange region because neighbours write their halo rows
enum DistField { DF_PSI = 0, DF_PSI_MIDDLE, DF_BY, DF_BZ, DF_JX, DF_COUNT };

struct DistLayout {  // byte offsets inside an exchange region (identical on all ranks)
    size_t flags;        // long long[DIST_MAX]: epoch of the last barrier each rank has entered
    size_t stops;        // int[2][DIST_MAX]: Ctl::stop_tag of each rank as of its last barrier, double buffered by the parity of the epoch
    size_t partials;     // double[DIST_MAX][4]: border-sum partials of the fused potential fix-up
    size_t inbox_count;  // int[2]
    size_t cols;         // double2[ny * cpr_max]: this rank's columns of the spectrum
    size_t rbuf;         // double[RB_COUNT][2][DIST_HALO * nz]
    size_t inbox;        // double[2][6][mig_cap]
    size_t field[DF_COUNT];  // double[ny * nz]
    size_t total;
};

struct DistView {
    int rank, world;
    int below, above;  // neighbour ranks (periodic)
    int ny_loc, row0;
    int mig_cap;
    char * base[DIST_MAX];
    DistLayout lay;
    long long * epoch;  // local, device
    int * error;        // local, device: != 0 after a bounded spin gave up or a buffer overflowed
};

// region of rank q (the table is a kernel parameter: index it with compile-time constants only, or every thread copies it
// to local memory)
__device__ __forceinline__ char * dist_base(const DistView & d, int q) {
    char * b = d.base[0];
#pragma unroll
    for (int i = 1; i < DIST_MAX; i++) {
        if (q == i) b = d.base[i];
    }
    return b;
}

__device__ __forceinline__ void st_release_sys(long long * p, long long v) {
#ifdef QPIC_EMU
    __atomic_store_n(p, v, __ATOMIC_RELEASE);
#else
    asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}

__device__ __forceinline__ long long ld_acquire_sys(const long long * p) {
#ifdef QPIC_EMU
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
    long long v;
    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
#endif
}

__device__ __forceinline__ unsigned long long dist_now_ns() {
#ifdef QPIC_EMU
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
#else
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
#endif
}

// Everything this GPU stored before the barrier (previous kernels of the stream) is visible to the peers once they leave
// it, and vice versa.  NOT skipped after a stop / abort: the peers would wait for ever.
// The |B| stop condition (src/System_3d.cpp:545-550) is made global here: every rank publishes its stop tag with the barrier
// and adopts a peer's, so that from the first barrier after the B solve that tripped it (the By / Bz halo fill) all ranks
// cancel the rest of the sweep at the same point of their streams, as the single-GPU sweep does.
__global__ void k_xchg_barrier(DistView d, Ctl * ctl) {
    QPIC_KERNEL_BEGIN(KID_k_xchg_barrier);
    __shared__ long long ep;
    if (threadIdx.x == 0) {
        ep = *d.epoch + 1;
        *d.epoch = ep;
    }
    __syncthreads();
    const int q = threadIdx.x;
    // (slot of this epoch's parity: a fast peer that is already in barrier ep + 1 writes the other slot, and it cannot reach ep + 2 before this
    // rank has arrived at ep + 1, i.e. has read its tag of epoch ep)
    const int par = (int)(ep & 1) * DIST_MAX;
    if (q < d.world) reinterpret_cast<int *>(dist_base(d, q) + d.lay.stops)[par + d.rank] = ctl->stop_tag;
    __threadfence_system();
    if (q < d.world) {
        st_release_sys(reinterpret_cast<long long *>(dist_base(d, q) + d.lay.flags) + d.rank, ep);
        const long long * mine = reinterpret_cast<const long long *>(dist_base(d, d.rank) + d.lay.flags) + q;
        const unsigned long long t0 = dist_now_ns();
        while (ld_acquire_sys(mine) < ep) {
            if (dist_now_ns() - t0 > 4000000000ull) {  // 4 s: a peer died or the ranks diverged
                atomicExch(d.error, 1);
                break;
            }
        }
        // rank q's tag was stored before its flag (release) and is read after it (acquire)
        const int peer_tag = reinterpret_cast<const int *>(dist_base(d, d.rank) + d.lay.stops)[par + q];
        if (peer_tag != 0) atomicCAS(&ctl->stop_tag, 0, peer_tag);
    }
    __threadfence_system();
}

// rows of a full-size (ny, nz) array, periodic in the row index
__device__ __forceinline__ int dist_row(int j, int ny) {
    j %= ny;
    return j < 0 ? j + ny : j;
}

struct HaloJob {
    int n;                  // arrays in this exchange
    double * a[4];          // reduce: local full-size arrays
    int slot[4];            // DistReduceSlot (reduce) / DistField (fill)
    int slice3d[4];         // != 0: a[] is a 3-D array, the job works on slice ctl->slice of it
    size_t field_off[4];    // fill: byte offset of the field inside an exchange region
};

// element i of a small kernel-parameter array without dynamic indexing (which would cost every thread a local-memory copy --
// and a first launch that needs local memory synchronises the device, fatal while a peer context spins in a barrier)
template <class T>
__device__ __forceinline__ T pick4(const T (&arr)[4], int i) {
    T v = arr[0];
#pragma unroll
    for (int t = 1; t < 4; t++) {
        if (i == t) v = arr[t];
    }
    return v;
}

// partial sums of the DIST_HALO rows below / above the slab -> the neighbours' reduce buffers
// rbuf[slot][0] receives from the rank above (their lower halo = my top rows), rbuf[slot][1] from the rank below
__global__ void k_halo_reduce_send(const Ctl * ctl, Geom g, DistView d, HaloJob job) {
    QPIC_KERNEL_BEGIN(KID_k_halo_reduce_send);
    const int per = DIST_HALO * g.nz;
    const int total = job.n * 2 * per;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int a = t / (2 * per), rem = t - a * 2 * per;
        const int side = rem / per, e = rem - side * per;  // side 0: my lower halo -> rank below; 1: my upper halo -> rank above
        const int row = side == 0 ? dist_row(d.row0 - DIST_HALO + e / g.nz, g.ny) : dist_row(d.row0 + d.ny_loc + e / g.nz, g.ny);
        const double * src = pick4(job.a, a) + (pick4(job.slice3d, a) ? (size_t)ctl->slice * g.n2 : 0);
        const double v = src[(size_t)row * g.nz + (e % g.nz)];
        char * peer = dist_base(d, side == 0 ? d.below : d.above);
        double * dst = reinterpret_cast<double *>(peer + d.lay.rbuf) + ((size_t)pick4(job.slot, a) * 2 + (side == 0 ? 0 : 1)) * per;
        dst[e] = v;
    }
}

// own edge rows += what the neighbours sent (fixed order: the two contributions go to different rows)
__global__ void k_halo_reduce_add(const Ctl * ctl, Geom g, DistView d, HaloJob job) {
    QPIC_KERNEL_BEGIN(KID_k_halo_reduce_add);
    const int per = DIST_HALO * g.nz;
    const int total = job.n * 2 * per;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int a = t / (2 * per), rem = t - a * 2 * per;
        const int side = rem / per, e = rem - side * per;  // side 0: from the rank above -> my top rows; 1: from below -> my first rows
        const int row = side == 0 ? d.row0 + d.ny_loc - DIST_HALO + e / g.nz : d.row0 + e / g.nz;
        double * dst = pick4(job.a, a) + (pick4(job.slice3d, a) ? (size_t)ctl->slice * g.n2 : 0);
        const double * src = reinterpret_cast<const double *>(dist_base(d, d.rank) + d.lay.rbuf) + ((size_t)pick4(job.slot, a) * 2 + side) * per;
        dst[(size_t)row * g.nz + (e % g.nz)] += src[e];
    }
}

// my first / last DIST_HALO rows of the region fields -> the same rows of the neighbours' copies (their upper / lower halo)
__global__ void k_halo_fill(Geom g, DistView d, HaloJob job) {
    QPIC_KERNEL_BEGIN(KID_k_halo_fill);
    const int per = DIST_HALO * g.nz;
    const int total = job.n * 2 * per;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int a = t / (2 * per), rem = t - a * 2 * per;
        const int side = rem / per, e = rem - side * per;  // side 0: my first rows -> rank below; 1: my last rows -> rank above
        const int row = side == 0 ? d.row0 + e / g.nz : d.row0 + d.ny_loc - DIST_HALO + e / g.nz;
        const size_t idx = (size_t)row * g.nz + (e % g.nz);
        const double * src = reinterpret_cast<const double *>(dist_base(d, d.rank) + pick4(job.field_off, a));
        double * dst = reinterpret_cast<double *>(dist_base(d, side == 0 ? d.below : d.above) + pick4(job.field_off, a));
        dst[idx] = src[idx];
    }
}

// ---- particle migration (after the push of a slice).  Emigrants were appended to the neighbours' inboxes by
// k_particles_advance; this publishes the counts, and after the barrier k_migrate_unpack appends the immigrants to the
// particle-id arrays (ids beyond the live slots), ready for the cell sort.
__global__ void k_migrate_publish(DistView d, int * out_count) {
    QPIC_KERNEL_BEGIN(KID_k_migrate_publish);
    if (threadIdx.x < 2 && blockIdx.x == 0) {
        const int side = threadIdx.x;  // 0: to the rank below (box 0 = "from above" there), 1: to the rank above (box 1)
        const int n = out_count[side];
        if (n > d.mig_cap) atomicExch(d.error, 2);
        int * peer = reinterpret_cast<int *>(dist_base(d, side == 0 ? d.below : d.above) + d.lay.inbox_count);
        peer[side] = n < d.mig_cap ? n : d.mig_cap;
        out_count[side] = 0;
    }
}

__global__ void k_migrate_unpack(Ctl * ctl, Geom g, DistView d, Particles P, int capacity, int * __restrict__ count) {
    QPIC_KERNEL_BEGIN(KID_k_migrate_unpack);
    const int * n_in = reinterpret_cast<const int *>(dist_base(d, d.rank) + d.lay.inbox_count);
    const int n0 = n_in[0], n1 = n_in[1];
    const int base = live_count(ctl, P.np);
    const double * box = reinterpret_cast<const double *>(dist_base(d, d.rank) + d.lay.inbox);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n0 + n1; t += gridDim.x * blockDim.x) {
        const int side = t < n0 ? 0 : 1, i = t < n0 ? t : t - n0;
        const int p = base + t;
        if (p >= capacity) {
            atomicExch(d.error, 3);
            continue;
        }
        const double * b = box + (size_t)side * 6 * d.mig_cap;
        const double y = b[0 * (size_t)d.mig_cap + i], z = b[1 * (size_t)d.mig_cap + i];
        P.y[p] = y;
        P.z[p] = z;
        P.py[p] = b[2 * (size_t)d.mig_cap + i];
        P.pz[p] = b[3 * (size_t)d.mig_cap + i];
        P.nw[p] = b[4 * (size_t)d.mig_cap + i];
        // canonical order of the arrivals inside a cell: (side, slot on the sending rank), see sort_key; slots < 2^30
        P.akey[t] = (side << 30) | (int)b[5 * (size_t)d.mig_cap + i];
        P.gamma[p] = 0.0;
        P.px[p] = 0.0;
        int jn, kn;
        double f1, f2;
        locate(y, g.dy, g.ny, jn, f1);
        locate(z, g.dz, g.nz, kn, f2);
        const int cn = jn * g.nz + kn;
        P.cell_p[p] = cn;
        P.fy_p[p] = f1;
        P.fz_p[p] = f2;
        atomicAdd(&count[cn], 1);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int total = base + n0 + n1;
        if (total > capacity) total = capacity;
        ctl->np_p1 = total + 1;
        ctl->mig_base1 = base + 1;
    }
}

}  // namespace qpic
