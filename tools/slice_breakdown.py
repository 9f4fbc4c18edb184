"""Per-kernel breakdown of a slice (eager, one CUDA event per launch) next to the graph-replay time (GPU box)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quetzalpic_b200 import WakefieldSolver, gaussian_laser_widths, wake_particle_lattice  # noqa: E402

n, nx, M = 512, 256, 5
l, dx = 32.0, 0.015625
s = WakefieldSolver(nx, n, n, nx * dx, l, l, M, force_general_deposit="--general" in sys.argv)
s.set_laser_gaussian(0.5, *gaussian_laser_widths(4.0, 8.0, 8.0), 0.5 * nx * dx, 0.5 * l, 0.5 * l)
s.init_wake_particles(*wake_particle_lattice(n, n, l, l, 2, 2))
s.begin()
s.step(32, wait=True)
s.step(96, wait=True)
ms, launches = s.last_step_stats()
print(f"graph replay: {1e3 * ms / 96:.1f} us/slice, {launches / 96:.0f} kernels/slice")
prof = s.profile_slices(16)
tot = sum(v[1] for v in prof.values())
print(f"eager, event after every kernel: {1e3 * tot / 16:.1f} us/slice")
for name, (cnt, t) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"  {name:32s} x{cnt / 16:5.1f}  {1e3 * t / cnt:7.2f} us  {1e3 * t / 16:8.1f} us/slice  {100 * t / tot:5.1f}%")
s.close()
