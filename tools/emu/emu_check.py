"""DEVELOPMENT TOOL: runs the kernels under the CPU fiber emulator (tools/emu) against the parity oracle.

Used in the GPU-less build container to debug kernel logic before spending GPU-box minutes.  The emulator library is
a separate debug artefact; the quetzalpic_b200 package never loads it.
Usage: tools/emu/build_emu.sh && python tools/emu/emu_check.py [case ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.qo import Oracle  # noqa: E402
from quetzalpic_b200.solver import QpicLibrary, WakefieldSolver, gaussian_laser_widths, wake_particle_lattice, powerlaw2d  # noqa: E402

EMU = QpicLibrary(os.path.join(ROOT, "tools", "emu", "_build", "libqpic_emu_debug.so"))

CASES = {
    # name: nx, ny, nz, lx, ly, lz, M, ppcy, ppcz, a0, (xs, ys, zs), x0
    "tiny_laser": dict(n=(16, 20, 23), l=(3.2, 4.0, 4.8), M=3, ppc=(2, 1), a0=1.5, sig=(1.2, 2.0, 2.4), x0=2.2),
    "pow2": dict(n=(6, 16, 32), l=(1.2, 3.2, 6.4), M=2, ppc=(2, 2), a0=1.2, sig=(0.8, 1.6, 3.0), x0=0.8),
    "odd": dict(n=(5, 15, 9), l=(1.0, 3.0, 1.8), M=2, ppc=(1, 2), a0=1.0, sig=(0.8, 1.6, 1.0), x0=0.6),
    "m0": dict(n=(4, 8, 12), l=(0.8, 1.6, 2.4), M=0, ppc=(1, 1), a0=1.0, sig=(0.8, 1.6, 1.0), x0=0.5),
}


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def run_case(name, cfg):
    nx, ny, nz = cfg["n"]
    lx, ly, lz = cfg["l"]
    M = cfg["M"]
    amp = 0.5 * cfg["a0"] ** 2
    wx, wy, wz = gaussian_laser_widths(*cfg["sig"])
    y, z, n = wake_particle_lattice(ny, nz, ly, lz, *cfg["ppc"])
    o = Oracle(nx, ny, nz, lx, ly, lz, M)
    o.set_laser_gaussian(amp, wx, wy, wz, cfg["x0"], 0.5 * ly, 0.5 * lz)
    o.init_wake_particles(y, z, n)
    t0 = time.time()
    o.solve_wakefield()
    t1 = time.time()
    s = WakefieldSolver(nx, ny, nz, lx, ly, lz, M, keep_slices_3d=True, library=EMU)
    s.upload_a_sqr(o.field3d("aSqr"))
    s.init_wake_particles(y, z, n)
    s.solve_wakefield()
    t2 = time.time()
    worst = 0.0
    for f in ["susceptibility", "ex", "ey", "ez", "by", "bz"]:
        e = rel(s.field3d(f), o.field3d(f))
        worst = max(worst, e)
        print(f"  {f:15s} {e:.3e}")
    for f in ["psi", "rho", "jx", "jy", "jz"]:
        e = rel(s.field3d(f, per_slice=True), o.field3d(f, per_slice=True))
        worst = max(worst, e)
        print(f"  {f:15s} {e:.3e}")
    for f in ["y", "z", "py", "pz", "gamma", "px"]:
        e = rel(s.particles(f), o.particles(f))
        worst = max(worst, e)
        print(f"  p.{f:13s} {e:.3e}")
    print(f"{name}: worst {worst:.3e}  (oracle {t1 - t0:.2f}s, emu {t2 - t1:.2f}s) stopped {s.stopped_at} / {o.stopped_at}")
    s.close()
    return worst


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    bad = 0
    for nm in names:
        if run_case(nm, CASES[nm]) > 1e-11:
            bad += 1
    sys.exit(bad)
