"""Times the beam ("bunch") steps on the GPU box: deposit_bunch_densities, deposit_wakefield_to_bunch, advance_bunch_particles
(reference src/System_3d.cpp:926-946, :843-860, :862-886) on the shape of the shipped inputs/test/bunch_test.toml
(200^3 cells, Gaussian beam filling the box = 8 M particles) and on a 1024 x 256 x 256 grid with a compact 16 M-particle beam.
Prints ms per call and the algorithmic HBM traffic rate.  python tools/beam_bench.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quetzalpic_b200 import WakefieldSolver, bunch_particle_lattice, gaussian3d, parabolic3d  # noqa: E402


def timed(fn, reps=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3


def run(name, n, l, particles):
    count = particles["x"].size
    n3 = n[0] * n[1] * n[2]
    with WakefieldSolver(*n, *l, 1) as s:
        s.bunch_init(particles)
        dep = timed(s.deposit_bunch_densities)
        gat = timed(lambda: s.deposit_wakefield_to_bunch(1.0))
        adv = timed(lambda: s.advance_bunch_particles(1e-6))
        # algorithmic bytes: deposit reads x y z px gamma n (+ int species, int cell key) and writes two 3-D arrays once;
        # gather reads x y z px py pz gamma, 5 x 8 field values (L2), writes 6; advance reads 12 + species, writes 7
        b_dep = count * (6 * 8 + 8) + 2 * 8 * n3
        b_gat = count * (7 * 8 + 6 * 8 + 5 * 8 * 8)
        b_adv = count * (12 * 8 + 4 + 7 * 8)
        print(f"{name}: grid {n}, {count / 1e6:.2f} M beam particles | deposit {dep:.3f} ms ({b_dep / dep / 1e6:.0f} GB/s alg.) | "
              f"gather {gat:.3f} ms ({b_gat / gat / 1e6:.0f} GB/s alg.) | advance {adv:.3f} ms ({b_adv / adv / 1e6:.0f} GB/s alg.)", flush=True)


def main():
    n, l = (200, 200, 200), (20.0, 20.0, 20.0)
    w = 4.0 / np.sqrt(8.0)
    run("bunch_test shape", n, l, bunch_particle_lattice(n, l, (1, 1, 1), gaussian3d(3.0, (w, w, w), (15.0, 10.0, 10.0)), 1000.0))
    n, l = (1024, 256, 256), (16.0, 16.0, 16.0)
    run("compact beam, 2x2x2 per cell", n, l, bunch_particle_lattice(n, l, (2, 2, 2), parabolic3d(2.0, 2.0, 2.0, (10.0, 8.0, 8.0)), 800.0))


if __name__ == "__main__":
    main()
