"""Slab-decomposed sweep with ONE PROCESS PER GPU (CUDA IPC exchange regions, torch.distributed for the handle exchange).

  parity:  torchrun --nproc-per-node 2 tools/dist_run.py --check
           every rank runs its slab of a small problem; rank 0 also runs the whole problem on its GPU alone and on the CPU
           oracle; the stitched fields must agree to 1e-11.
  bench:   torchrun --nproc-per-node N tools/dist_run.py --bench --n 4096 --ppc 4 --nx 64 [--slices 24]
           BASELINE config #4 shape (4096 x 4096 transverse grid, 16 ppc, M = 5); prints one JSON line (rank 0).
           N = 1 (plain `python tools/dist_run.py --bench ...`) runs the same grid on a single GPU for the scaling baseline.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def lattice_rows(ny, nz, ly, lz, ppcy, ppcz, row0, nrows, rho0=-1.0):
    """The rows [row0, row0 + nrows) of System_3d::init_wake_particles' lattice (src/System_3d.cpp:724-748) without building
    the whole (268 M particle) lattice on every rank."""
    dy, dz = ly / ny, lz / nz
    yy = (np.arange(ppcy * row0, ppcy * (row0 + nrows)) + 0.5) * dy / ppcy
    zz = (np.arange(ppcz * nz) + 0.5) * dz / ppcz
    y, z = np.meshgrid(yy, zz, indexing="ij")
    y, z = y.ravel(), z.ravel()
    return y, z, np.full(y.shape, float(rho0) / ppcy / ppcz)


def make_rank_solver(nx, ny, nz, lx, ly, lz, M, laser, ppc, rank, world, local, **kw):
    import torch.distributed as dist
    from quetzalpic_b200 import WakefieldSolver
    from quetzalpic_b200.dist import attach_process_group, slab_rows
    s = WakefieldSolver(nx, ny, nz, lx, ly, lz, M, device=local, dist_rank=rank, dist_world=world, **kw)
    if world > 1:
        attach_process_group(s)
    s.set_laser_gaussian(*laser)
    r0, nr = slab_rows(ny, rank, world) if world > 1 else (0, ny)
    s.init_wake_particles(*lattice_rows(ny, nz, ly, lz, ppc[0], ppc[1], r0, nr))
    if world > 1:
        dist.barrier()
    return s


def check(args):
    import torch
    import torch.distributed as dist
    from cases import Case, slice_rel_l2
    from quetzalpic_b200.dist import slab_rows
    rank, world, local = env()
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    n = (10, 64 * world, 128)
    case = Case.synthetic(n=n, d=(0.2, 0.1, 0.1), M=3, ppc=(2, 2), a0=1.8, sigma=(1.2, 1.6 * world, 3.0))
    wx, wy, wz = case.laser[1:4]
    laser = (case.laser[0], wx, wy, wz, 1.2, 0.1 * 64 + 0.05, 6.4)  # on the edge between ranks 0 and 1
    s = make_rank_solver(n[0], n[1], n[2], case.l[0], case.l[1], case.l[2], case.M, laser, (2, 2), rank, world, local, stop_check_interval=4)
    s.solve_wakefield()
    counts = [None] * world
    dist.all_gather_object(counts, int(s.particle_count))
    fields = ["susceptibility", "ex", "ey", "ez", "by", "bz"]
    r0, nr = slab_rows(n[1], rank, world)
    mine = {f: s.field3d(f)[:, r0:r0 + nr].copy() for f in fields}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    s.close()
    ok = True
    if rank == 0:
        case.laser = laser
        o = case.make_oracle()
        o.solve_wakefield()
        with case.make_solver(device=local) as single:
            single.solve_wakefield()
            worst = 0.0
            for f in fields:
                got = np.concatenate([g[f] for g in gathered], axis=1)
                e1, e2 = slice_rel_l2(got, o.field3d(f)), slice_rel_l2(got, single.field3d(f))
                worst = max(worst, e1, e2)
                print(f"  {f:16s} vs oracle {e1:.2e}   vs single GPU {e2:.2e}")
        initial = case.particles[0].size
        print(f"dist check: world={world}, particles per rank {counts} (sum {sum(counts)} of {initial}), worst error {worst:.2e}")
        ok = worst < 1e-11 and sum(counts) == initial
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


def bench(args):
    import torch
    import torch.distributed as dist
    rank, world, local = env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    n, nx, M = args.n, args.nx, args.M
    d = 64.0 / 4096  # BASELINE config #4: ly = lz = 64 on 4096 cells; dx the same
    lx, l = nx * d, n * d
    from quetzalpic_b200 import gaussian_laser_widths
    wx, wy, wz = gaussian_laser_widths(0.5, 0.25 * l, 0.25 * l)
    laser = (0.5 * args.a0 ** 2, wx, wy, wz, 0.6 * lx, 0.5 * l, 0.5 * l)
    t0 = time.time()
    s = make_rank_solver(nx, n, n, lx, l, l, M, laser, (args.ppc, args.ppc), rank, world, local, stop_check_interval=10 ** 6)
    P_total = float(n) * n * args.ppc * args.ppc
    setup_s = time.time() - t0
    s.begin()
    s.step(args.warm, wait=False)
    s.synchronize()
    if world > 1:
        dist.barrier()
    s.step(args.slices, wait=False)
    s.synchronize()
    ms, launches = s.last_step_stats()
    t = torch.tensor([ms], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    live = int(s.particle_count) if world > 1 else int(P_total)
    counts = [live]
    if world > 1:
        counts = [None] * world
        dist.all_gather_object(counts, live)
    free, total = torch.cuda.mem_get_info(local)
    if rank == 0:
        N = float(n) * n
        alg = 8.0 * ((29 + 44 * M) * N + (18 + 21 * M) * P_total)
        line = {
            "metric": "macroparticle pushes/sec (3D laser wake, slab-decomposed transverse grid)", "value": P_total * args.slices / (ms_max * 1e-3),
            "unit": "pushes/s", "n_gpus": world, "ms_per_slice": ms_max / args.slices, "slices": args.slices, "warmup_slices": args.warm,
            "scaling": "strong", "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"laser_wake_{n}x{n}_{args.ppc * args.ppc}ppc_M{M}_slab", "grid": [nx, n, n], "particles": P_total,
                       "parallelism": f"y-slabs x{world}, peer-store transposes + halo exchange over NVLink" if world > 1 else "single GPU"},
            "roofline_slice": {"bound": "hbm", "algorithmic_bytes_per_slice": alg, "achieved_gbs_aggregate": alg / (ms_max / args.slices * 1e-3) / 1e9,
                               "achieved_gbs_per_gpu": alg / (ms_max / args.slices * 1e-3) / 1e9 / world},
            "particles_per_rank_after": counts, "kernels_per_slice": launches // args.slices, "setup_s": setup_s,
            "hbm_used_gb_rank0": (total - free) / 1e9,
        }
        print(json.dumps(line), flush=True)
    s.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--nx", type=int, default=64)
    ap.add_argument("--ppc", type=int, default=4, help="per axis (4 -> 16 ppc)")
    ap.add_argument("--M", type=int, default=5)
    ap.add_argument("--a0", type=float, default=1.0)
    ap.add_argument("--slices", type=int, default=24)
    ap.add_argument("--warm", type=int, default=4)
    args = ap.parse_args()
    if args.check:
        return check(args)
    return bench(args)


if __name__ == "__main__":
    sys.exit(main())
