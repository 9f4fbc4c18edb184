import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
from quetzalpic_b200 import WakefieldSolver, gaussian_laser_widths, load_library, wake_particle_lattice
lib = load_library()
nx, n, M, S = 1024, 512, 5, 64
l, dx = 32.0, 0.015625
s = WakefieldSolver(nx, n, n, nx*dx, l, l, M, stop_check_interval=S)
s.set_laser_gaussian(0.5, *gaussian_laser_widths(4.0, 8.0, 8.0), 12.0, 16.0, 16.0)
y, z, w = wake_particle_lattice(n, n, l, l, 2, 2)
s.init_wake_particles(y, z, w); s.begin(); s.step(64, wait=True)
bufs = [lib.pinned_array((S, n, n))[0] for _ in range(6)]
names = ["ex", "ey", "ez", "by", "bz", "susceptibility"]
# (a) raw D2H bandwidth through the library's copy stream
for rep in range(2):
    t0 = time.perf_counter()
    for b, f in zip(bufs, names): s.download_slices3d(f, 100, b)
    s.wait_transfers()
    t1 = time.perf_counter()
    print(f"D2H 6 x {bufs[0].nbytes/1e6:.0f} MB: {t1-t0:.4f} s = {6*bufs[0].nbytes/1e9/(t1-t0):.1f} GB/s")
# torch pinned for comparison
a = torch.empty(S*n*n, dtype=torch.float64, device='cuda'); h = torch.empty(S*n*n, dtype=torch.float64).pin_memory()
torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(a, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"torch pinned D2H {h.numel()*8/1e6:.0f} MB: {h.numel()*8/1e9/(t1-t0):.1f} GB/s")
# (b) compute-only step vs step overlapped with download
s.synchronize()
t0 = time.perf_counter(); s.step(S, wait=False); s.synchronize(); t1 = time.perf_counter()
print(f"step alone: {t1-t0:.4f} s")
t0 = time.perf_counter(); s.step(S, wait=False)
for b, f in zip(bufs, names): s.download_slices3d(f, 100, b)
s.step(S, wait=False); s.synchronize(); t1 = time.perf_counter()
print(f"2 steps with a download in between (overlapped): {t1-t0:.4f} s")
ms2, _ = s.last_step_stats()
print(f"device time of the second step (events): {ms2:.1f} ms")
s.close()
for graph in (True, False):
    s = WakefieldSolver(nx, n, n, nx*dx, l, l, M, stop_check_interval=10**6, use_graph=graph)
    s.set_laser_gaussian(0.5, *gaussian_laser_widths(4.0, 8.0, 8.0), 12.0, 16.0, 16.0)
    s.init_wake_particles(y, z, w); s.begin(); s.step(64, wait=False); s.lib.L.qpic_synchronize(s.ctx)
    t0 = time.perf_counter(); s.step(S, wait=False); s.lib.L.qpic_wait_transfers(s.ctx); torch.cuda.synchronize(); s.synchronize(); t1 = time.perf_counter()
    print(f"graph={graph}: step alone {t1-t0:.4f} s")
    t0 = time.perf_counter()
    for b, f in zip(bufs, names): s.download_slices3d(f, 100, b)
    s.step(S, wait=False); s.synchronize(); t1 = time.perf_counter()
    print(f"graph={graph}: download + step queued together, no polling: {t1-t0:.4f} s")
    s.close()
