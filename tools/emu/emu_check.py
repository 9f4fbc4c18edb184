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

FORCE_GENERAL = False

CASES = {
    # strong drive on a coarse longitudinal step: half-step displacements exceed one cell -> fast path must roll back
    "abort": dict(n=(10, 16, 16), l=(4.0, 3.2, 3.2), M=2, ppc=(1, 1), a0=2.5, sig=(2.0, 1.6, 1.6), x0=2.8, poll=3),
    # name: nx, ny, nz, lx, ly, lz, M, ppcy, ppcz, a0, (xs, ys, zs), x0
    "tiny_laser": dict(n=(16, 20, 23), l=(3.2, 4.0, 4.8), M=3, ppc=(2, 1), a0=1.5, sig=(1.2, 2.0, 2.4), x0=2.2),
    "pow2": dict(n=(6, 16, 32), l=(1.2, 3.2, 6.4), M=2, ppc=(2, 2), a0=1.2, sig=(0.8, 1.6, 3.0), x0=0.8),
    "odd": dict(n=(5, 15, 9), l=(1.0, 3.0, 1.8), M=2, ppc=(1, 2), a0=1.0, sig=(0.8, 1.6, 1.0), x0=0.6),
    "p64": dict(n=(4, 64, 128), l=(0.8, 6.4, 12.8), M=2, ppc=(1, 1), a0=1.3, sig=(0.8, 3.0, 6.0), x0=0.5),
    "p256": dict(n=(3, 256, 64), l=(0.6, 25.6, 6.4), M=1, ppc=(1, 1), a0=1.3, sig=(0.8, 12.0, 3.0), x0=0.4),
    # power-of-two steps (dy = dz = 1/16): the kernels rebuild positions from the cell descriptor and locate cells by multiplication
    "exact": dict(n=(6, 32, 48), l=(0.75, 2.0, 3.0), M=3, ppc=(2, 2), a0=2.0, sig=(0.6, 0.8, 1.0), x0=0.4),
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
    s = WakefieldSolver(nx, ny, nz, lx, ly, lz, M, keep_slices_3d=True, library=EMU, force_general_deposit=FORCE_GENERAL,
                        stop_check_interval=cfg.get("poll", 64))
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
    print(f"{name}: worst {worst:.3e}  (oracle {t1 - t0:.2f}s, emu {t2 - t1:.2f}s) stopped {s.stopped_at} / {o.stopped_at}"
          f"  fast-path aborts {s.fast_path_aborts}  general={FORCE_GENERAL}")
    s.close()
    return worst


def fft_check():
    from oracle.qo import lib
    bad = 0
    for ny, nz in [(64, 128), (128, 64), (256, 64), (64, 1024), (20, 23), (512, 64)]:
        rng = np.random.default_rng(ny + nz)
        src, src2 = rng.standard_normal((ny, nz)), rng.standard_normal((ny, nz))
        o = Oracle(1, ny, nz, 1.0, 0.1 * ny, 0.07 * nz, 0)
        s = WakefieldSolver(1, ny, nz, 1.0, 0.1 * ny, 0.07 * nz, 0, library=EMU)
        for D in (0.0, 10.0):
            ref, ref2 = src.copy(), src2.copy()
            lib().qo_solve_poisson(o.ctx, ref.ctypes.data, D)
            lib().qo_solve_poisson(o.ctx, ref2.ctypes.data, D)
            ref /= ny * nz
            ref2 /= ny * nz
            e1 = rel(s.solve_poisson(src, D), ref)
            a, b = s.solve_helmholtz_pair(src, src2, D)
            e2, e3 = rel(a, ref), rel(b, ref2)
            print(f"fft {ny}x{nz} D={D}: real {e1:.2e} pair {e2:.2e} {e3:.2e}")
            bad += (max(e1, e2, e3) > 1e-13)
        s.close()
    return bad


if __name__ == "__main__":
    if sys.argv[1:] == ["fft"]:
        sys.exit(int(fft_check()))
    if "--general" in sys.argv:
        FORCE_GENERAL = True
        sys.argv.remove("--general")
    names = sys.argv[1:] or list(CASES)
    bad = 0
    for nm in names:
        if run_case(nm, CASES[nm]) > 1e-11:
            bad += 1
    sys.exit(bad)
