"""DEVELOPMENT TOOL: the beam ("bunch") kernels under the CPU fiber emulator against the golden fixture of the unmodified
reference (tests/golden/tiny_beam.npz).  Usage: tools/emu/build_emu.sh && python tools/emu/emu_beam_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import Case, rel_l2, slice_rel_l2  # noqa: E402
from quetzalpic_b200.solver import QpicLibrary  # noqa: E402

EMU = QpicLibrary(os.path.join(ROOT, "tools", "emu", "_build", "libqpic_emu_debug.so"))


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "tiny_beam.npz"))
    case = Case.from_toml("tiny_beam")
    particles = case.bunch_particles()
    worst = [0.0]

    def make_backend():
        s = case.make_solver(keep_slices_3d=True, library=EMU)
        s.bunch_init(particles)
        return s, s

    def check(ti, s, beam):
        for f in ["rho_bunch", "jx_bunch", "ex", "ey", "ez", "by", "bz", "susceptibility"]:
            e = slice_rel_l2(s.field3d(f), g[f"it{ti}/f3/{f}"])
            worst[0] = max(worst[0], e)
            print(ti, f, e)
        for f in ["psi", "rho", "jx", "jy", "jz"]:
            e = slice_rel_l2(s.field3d(f, per_slice=True), g[f"it{ti}/f3/{f}"])
            worst[0] = max(worst[0], e)
            print(ti, f, e)
        for f in ["x", "y", "z", "px", "py", "pz", "gamma", "ex", "ey", "ez", "by", "bz", "chi"]:
            e = rel_l2(s.bunch(f), g[f"it{ti}/bunch/{f}"])
            worst[0] = max(worst[0], e)
            print(ti, "bunch", f, e)

    case.run_time_loop(make_backend, check)
    print("worst", worst[0])
    assert worst[0] < 1e-10


if __name__ == "__main__":
    main()
