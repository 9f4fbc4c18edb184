"""DEVELOPMENT TOOL: the slab-decomposed sweep (2 or 4 ranks = host threads) under the CPU fiber emulator vs the oracle.

Usage: tools/emu/build_emu.sh && python tools/emu/emu_dist_check.py [world]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.qo import Oracle  # noqa: E402
from quetzalpic_b200.dist import SlabGroup  # noqa: E402
from quetzalpic_b200.solver import QpicLibrary, gaussian_laser_widths, wake_particle_lattice  # noqa: E402

EMU = QpicLibrary(os.path.join(ROOT, "tools", "emu", "_build", "libqpic_emu_debug.so"))


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    nx, ny, nz = int(os.environ.get("NX", "5")), 64 * (world // 2), 64
    lx, ly, lz = float(os.environ.get("DX", "0.2")) * nx, 6.4 * (world // 2), 6.4
    M = 2
    a0 = float(os.environ.get("A0", "1.6"))
    wx, wy, wz = gaussian_laser_widths(0.8, 3.0, 3.0)
    y, z, n = wake_particle_lattice(ny, nz, ly, lz, 2, 1)
    o = Oracle(nx, ny, nz, lx, ly, lz, M)
    o.set_laser_gaussian(0.5 * a0 * a0, wx, wy, wz, 0.6, 0.5 * ly - 0.3 * ly / world, 0.5 * lz)
    o.init_wake_particles(y, z, n)
    o.solve_wakefield()
    with SlabGroup(world, nx, ny, nz, lx, ly, lz, M, library=EMU, stop_check_interval=2) as g:
        g.upload_a_sqr(o.field3d("aSqr"))
        g.init_wake_particles(y, z, n)
        before = [s.particle_count for s in g.ranks]
        stopped = g.solve_wakefield()
        print('  particles per rank', before, '->', [s.particle_count for s in g.ranks], ' fast-path aborts', [s.fast_path_aborts for s in g.ranks])
        worst = 0.0
        for f in ["susceptibility", "ex", "ey", "ez", "by", "bz"]:
            a, b = g.field3d(f), o.field3d(f)
            scale = max(np.linalg.norm(b[i]) for i in range(nx)) or 1.0
            err = max(np.linalg.norm(a[i] - b[i]) for i in range(nx)) / scale
            worst = max(worst, err)
            print(f"  {f:16s} {err:.3e}")
    print(f"world={world}: worst {worst:.3e}, stopped {stopped}")
    return 0 if worst < 1e-10 else 1


if __name__ == "__main__":
    sys.exit(main())
