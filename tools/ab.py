"""A/B timing helper (GPU box): graph-replayed slices of a BASELINE shape, device-timed (CUDA events on the solver's stream).
  QPIC_PREDICT_DEPOSIT=0 python tools/ab.py linear512 --slices 512 --reps 3
Prints one line per repetition: us/slice, kernels/slice.  Environment variables select the variant."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quetzalpic_b200 import WakefieldSolver, gaussian_laser_widths, wake_particle_lattice  # noqa: E402
from quetzalpic_b200.solver import QpicLibrary  # noqa: E402

SHAPES = {
    # name: (nx, ny, nz, lx, ly, lz, ppc, M, a0, sigmas, x0)
    "linear512": (2048, 512, 512, 32.0, 32.0, 32.0, 2, 5, 1.0, (4.0, 8.0, 8.0), 26.0),
    "bubble1024": (1024, 1024, 1024, 16.0, 32.0, 32.0, 2, 10, 4.0, (4.0, 8.0, 8.0), 12.0),
    "wide4096": (40, 4096, 4096, 0.625, 64.0, 64.0, 4, 5, 1.0, (0.5, 16.0, 16.0), 0.375),
}
ap = argparse.ArgumentParser()
ap.add_argument("shape")
ap.add_argument("--slices", type=int, default=256)
ap.add_argument("--warm", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--tag", default="")
ap.add_argument("--lib", default=None, help="alternative build of the library (tools/variants/*.so, see tools/build_variant.sh)")
args = ap.parse_args()
lib = QpicLibrary(args.lib) if args.lib else None
nx, ny, nz, lx, ly, lz, ppc, M, a0, sig, x0 = SHAPES[args.shape]
s = WakefieldSolver(nx, ny, nz, lx, ly, lz, M, stop_check_interval=10 ** 6, library=lib)
s.set_laser_gaussian(0.5 * a0 * a0, *gaussian_laser_widths(*sig), x0, 0.5 * ly, 0.5 * lz)
s.init_wake_particles(*wake_particle_lattice(ny, nz, ly, lz, ppc, ppc))
s.begin()
s.step(args.warm, wait=False)
s.synchronize()
out = []
for r in range(args.reps):
    s.step(args.slices, wait=False)
    s.synchronize()
    ms, launches = s.last_step_stats()
    out.append(1e3 * ms / args.slices)
env = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("QPIC_"))
print(f"{args.shape} {args.tag or (os.path.basename(args.lib) if args.lib else '')} [{env}]: " + " ".join(f"{v:.1f}" for v in out) + f" us/slice, {launches // args.slices} kernels/slice, aborts {s.fast_path_aborts}, stopped {s.stopped_at}", flush=True)
s.close()
