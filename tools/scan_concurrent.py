"""Independent sweeps sharing ONE GPU (BASELINE config #5, the a0 / density scan): a single sweep is a chain of latency-bound kernels
that leaves most of the GPU idle most of the time, so members of a parameter scan can be interleaved on the same device.  Runs
`members` sweeps of the config #2 shape concurrently (one context + one host thread each) and prints the aggregate pushes/s.

Usage (GPU box): python tools/scan_concurrent.py [--members 1 2 3 4] [--slices 256]
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quetzalpic_b200 import WakefieldSolver, gaussian_laser_widths, wake_particle_lattice  # noqa: E402

SCAN = [(0.5, -1.0), (1.0, -1.0), (2.0, -1.0), (3.0, -1.0), (0.5, -0.5), (1.0, -0.5), (2.0, -0.5), (3.0, -0.5)]


def run(members, slices, n=512, nx=640, M=5, ppc=2):
    l, dx = 32.0, 0.015625
    wx, wy, wz = gaussian_laser_widths(4.0, 8.0, 8.0)
    solvers = []
    for m in range(members):
        a0, rho0 = SCAN[m % len(SCAN)]
        s = WakefieldSolver(nx, n, n, nx * dx, l, l, M, stop_check_interval=10 ** 6)
        s.set_laser_gaussian(0.5 * a0 * a0, wx, wy, wz, nx * dx - 6.0, 0.5 * l, 0.5 * l)
        s.init_wake_particles(*wake_particle_lattice(n, n, l, l, ppc, ppc, rho0=rho0))
        s.begin()
        s.step(64, wait=True)
        solvers.append(s)
    start = threading.Barrier(members + 1)
    done = threading.Barrier(members + 1)

    def work(s):
        start.wait()
        s.step(slices, wait=False)
        s.synchronize()
        done.wait()

    threads = [threading.Thread(target=work, args=(s,)) for s in solvers]
    for t in threads:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    done.wait()
    t1 = time.perf_counter()
    for t in threads:
        t.join()
    P = ppc * ppc * n * n
    out = {"members_on_one_gpu": members, "slices_each": slices, "seconds": t1 - t0, "ms_per_slice_per_member": 1e3 * (t1 - t0) / slices,
           "aggregate_pushes_per_s": members * P * slices / (t1 - t0), "stopped": [s.stopped_at for s in solvers]}
    for s in solvers:
        s.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, nargs="+", default=[1, 2, 3, 4])
    ap.add_argument("--slices", type=int, default=256)
    a = ap.parse_args()
    for m in a.members:
        print(json.dumps(run(m, a.slices)), flush=True)
