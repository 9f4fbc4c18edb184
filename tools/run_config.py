"""Runs a few slices of one of the BASELINE.json shapes and prints ms/slice + fast-path aborts (GPU box).
  python tools/run_config.py bubble1024 --slices 96     # config #3: 1024^2, 4 ppc, a0 = 4, M = 10 (trajectory crossing)
  python tools/run_config.py laser_test                  # config #1: the shipped 200^3 input, whole sweep
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quetzalpic_b200 import WakefieldSolver, gaussian_laser_widths, wake_particle_lattice  # noqa: E402

SHAPES = {
    # name: (nx, ny, nz, lx, ly, lz, ppc, M, a0, sigmas, x0)
    "laser_test": (200, 200, 200, 20.0, 20.0, 20.0, 1, 5, 1.41421356237, (4.0, 4.0, 4.0), 15.0),
    "linear512": (2048, 512, 512, 32.0, 32.0, 32.0, 2, 5, 1.0, (4.0, 8.0, 8.0), 26.0),
    "bubble1024": (1024, 1024, 1024, 16.0, 32.0, 32.0, 2, 10, 4.0, (4.0, 8.0, 8.0), 12.0),
}
ap = argparse.ArgumentParser()
ap.add_argument("shape")
ap.add_argument("--slices", type=int, default=0)
ap.add_argument("--general", action="store_true")
args = ap.parse_args()
nx, ny, nz, lx, ly, lz, ppc, M, a0, sig, x0 = SHAPES[args.shape]
s = WakefieldSolver(nx, ny, nz, lx, ly, lz, M, force_general_deposit=args.general)
s.set_laser_gaussian(0.5 * a0 * a0, *gaussian_laser_widths(*sig), x0, 0.5 * ly, 0.5 * lz)
s.init_wake_particles(*wake_particle_lattice(ny, nz, ly, lz, ppc, ppc))
s.begin()
todo = args.slices or nx
done_slices = 0
t0 = time.perf_counter()
while done_slices < todo:
    n = min(64, todo - done_slices)
    before = s.current_slice
    finished = s.step(n, wait=True)
    done_slices += n
    if finished:
        break
t1 = time.perf_counter()
pushed = nx - 1 - max(s.current_slice, 0) if s.stopped_at < 0 else nx - 1 - s.stopped_at
print(f"{args.shape}: {pushed} slices in {t1 - t0:.3f} s = {1e3 * (t1 - t0) / max(pushed, 1):.3f} ms/slice, "
      f"{ppc * ppc * ny * nz * pushed / (t1 - t0) / 1e6:.1f} M pushes/s, stopped_at={s.stopped_at}, fast-path aborts={s.fast_path_aborts}, "
      f"general={args.general}")
s.close()
