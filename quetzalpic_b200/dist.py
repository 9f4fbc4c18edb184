"""Host side of the slab-decomposed multi-GPU sweep (BASELINE config #4).

One large transverse grid is cut into `world` slabs of ny/world rows, one WakefieldSolver context per GPU.  The data path
between GPUs is inside the library (peer loads / stores over NVLink in the slice graph, csrc/qpic_dist.cuh); the host only
  * deals the plasma lattice out by rows (`slab_particles`),
  * lets the ranks map each other's exchange regions -- between processes with CUDA IPC handles all-gathered through
    torch.distributed (`attach_process_group`), inside one process with raw pointers (`SlabGroup`, used by the tests and
    by single-process multi-GPU drivers),
  * and stitches per-rank field rows back together for output / comparison (`gather_rows`).
"""
import threading

import numpy as np

from .solver import WakefieldSolver


def slab_rows(ny, rank, world):
    """(first row, number of rows) of a rank: rows are dealt out in equal contiguous blocks."""
    if ny % world != 0:
        raise ValueError("ny must be divisible by the number of ranks")
    n = ny // world
    return rank * n, n


def half_spectrum_columns(nz, rank, world):
    """(first column, count) of the nz/2+1 half-spectrum columns a rank transforms (the last ranks may hold fewer)."""
    nzh = nz // 2 + 1
    cpr = (nzh + world - 1) // world
    c0 = rank * cpr
    return c0, max(0, min(cpr, nzh - c0))


def slab_particles(y, z, n, ny, ly, rank, world):
    """The particles whose cell row (int)(y / dy) lies in the rank's slab (the lattice order is kept)."""
    dy = ly / ny
    row = (y / dy).astype(np.int64)
    row[row >= ny] = 0  # same convention as the device code for y within an ulp of ly
    r0, nr = slab_rows(ny, rank, world)
    keep = (row >= r0) & (row < r0 + nr)
    return y[keep].copy(), z[keep].copy(), n[keep].copy()


def attach_process_group(solver, group=None):
    """One process per GPU: all-gather the CUDA IPC handles of the exchange regions and map the peers."""
    import torch
    import torch.distributed as dist

    mine = torch.frombuffer(bytearray(solver.dist_ipc_handle()), dtype=torch.uint8).clone()
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    if backend == "nccl":
        mine = mine.cuda()
    bucket = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bucket, mine, group=group)
    solver.dist_attach_ipc([bytes(t.cpu().numpy().tobytes()) for t in bucket])
    dist.barrier(group=group)


def gather_rows(per_rank_arrays, ny, axis):
    """Own rows of every rank's (full-size) array, stitched together along `axis`."""
    world = len(per_rank_arrays)
    out = np.empty_like(per_rank_arrays[0])
    for r, a in enumerate(per_rank_arrays):
        r0, nr = slab_rows(ny, r, world)
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(r0, r0 + nr)
        out[tuple(sl)] = a[tuple(sl)]
    return out


class SlabGroup:
    """`world` slab contexts driven from one process (one host thread per rank while the sweep runs).

    devices: CUDA device ordinal per rank; the same ordinal may repeat (the ranks then share a GPU -- how the parity tests
    exercise the exchange path on a single-GPU box).
    """

    def __init__(self, world, nx, ny, nz, lx, ly, lz, M, devices=None, library=None, **kw):
        self.world, self.nx, self.ny, self.nz = world, nx, ny, nz
        self.ly = ly
        devices = list(devices) if devices is not None else [0] * world
        # Ranks of ONE process share a CUDA context per device, and the first launch of a kernel that needs (more) local
        # memory makes the driver synchronise the device -- which never returns while a peer's barrier kernel is spinning.
        # Run every kernel of the slice once on a throw-away single-GPU context per device first.
        for dev in sorted(set(devices)):
            with WakefieldSolver(3, 64, 64, 0.6, 6.4, 6.4, max(M, 1), device=dev, library=library, use_graph=False) as warm:
                warm.set_laser_gaussian(0.5, 0.3, 1.0, 1.0, 0.4, 3.2, 3.2)
                from .solver import wake_particle_lattice
                warm.init_wake_particles(*wake_particle_lattice(64, 64, 6.4, 6.4, 1, 1))
                warm.solve_wakefield()
        self.ranks = [WakefieldSolver(nx, ny, nz, lx, ly, lz, M, device=devices[r], dist_rank=r, dist_world=world, library=library, **kw)
                      for r in range(world)]
        regions = [s.dist_region() for s in self.ranks]
        for s in self.ranks:
            s.dist_attach_pointers(regions)

    def each(self, fn):
        """fn(rank, solver) on every rank concurrently (the exchange barriers need all ranks to make progress)."""
        errors = [None] * self.world

        def run(r):
            try:
                fn(r, self.ranks[r])
            except BaseException as e:  # noqa: BLE001 -- re-raised below
                errors[r] = e

        threads = [threading.Thread(target=run, args=(r,)) for r in range(self.world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in errors:
            if e is not None:
                raise e

    def set_laser_gaussian(self, *args):
        for s in self.ranks:
            s.set_laser_gaussian(*args)

    def upload_a_sqr(self, a_sqr):
        for s in self.ranks:
            s.upload_a_sqr(a_sqr)

    def init_wake_particles(self, y, z, n):
        for r, s in enumerate(self.ranks):
            s.init_wake_particles(*slab_particles(y, z, n, self.ny, self.ly, r, self.world))

    def solve_wakefield(self):
        self.each(lambda r, s: s.solve_wakefield())
        return [s.stopped_at for s in self.ranks]

    def begin(self):
        self.each(lambda r, s: s.begin())

    def step(self, nslices):
        self.each(lambda r, s: s.step(nslices, wait=True))

    def field3d(self, name):
        return gather_rows([s.field3d(name) for s in self.ranks], self.ny, axis=1)

    def field2d(self, name):
        return gather_rows([s.field2d(name) for s in self.ranks], self.ny, axis=0)

    def close(self):
        for s in self.ranks:
            s.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
