"""Runs a few xi-slices of the bench workload's transverse shape (512 x 512, 4 ppc, M = 5) -- short enough to sit under
ncu.  Usage (on the GPU box):
  ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv --log-file gpurun_out/launches.csv \
      python tools/profile_slices.py [--slices 12] [--eager]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quetzalpic_b200 import WakefieldSolver, gaussian_laser_widths, wake_particle_lattice  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--slices", type=int, default=12)
ap.add_argument("--nx", type=int, default=64)
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--ppc", type=int, default=2)
ap.add_argument("--M", type=int, default=5)
ap.add_argument("--eager", action="store_true")
args = ap.parse_args()
n = args.n
l = 32.0 * n / 512
dx = 0.015625
s = WakefieldSolver(args.nx, n, n, args.nx * dx, l, l, args.M, use_graph=not args.eager)
wx, wy, wz = gaussian_laser_widths(4.0, 8.0, 8.0)
s.set_laser_gaussian(0.5, wx, wy, wz, 0.5 * args.nx * dx, 0.5 * l, 0.5 * l)
s.init_wake_particles(*wake_particle_lattice(n, n, l, l, args.ppc, args.ppc))
s.begin()
s.step(args.slices, wait=True)
ms, launches = s.last_step_stats()
print(f"{args.slices} slices: {ms / args.slices:.4f} ms/slice, {launches} launches")
s.close()
