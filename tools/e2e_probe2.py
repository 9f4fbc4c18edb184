import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
from quetzalpic_b200 import WakefieldSolver, gaussian_laser_widths, load_library, wake_particle_lattice
nx, n, M, S = 512, 512, 5, 64
l, dx = 32.0, 0.015625
s = WakefieldSolver(nx, n, n, nx*dx, l, l, M, stop_check_interval=10**6)
s.set_laser_gaussian(0.5, *gaussian_laser_widths(4.0, 8.0, 8.0), 6.0, 16.0, 16.0)
y, z, w = wake_particle_lattice(n, n, l, l, 2, 2)
s.init_wake_particles(y, z, w); s.begin(); s.step(64, wait=False); s.synchronize()
d = torch.empty(100*1024*1024, dtype=torch.float64, device='cuda'); h = torch.empty_like(d, device='cpu').pin_memory()
cs = torch.cuda.Stream()
def tcopy():
    with torch.cuda.stream(cs): h.copy_(d, non_blocking=True)
tcopy(); torch.cuda.synchronize()
t0 = time.perf_counter(); tcopy(); torch.cuda.synchronize(); print("torch copy alone", round(time.perf_counter()-t0, 4))
t0 = time.perf_counter(); s.step(S, wait=False); s.synchronize(); print("step alone", round(time.perf_counter()-t0, 4))
t0 = time.perf_counter(); tcopy(); s.step(S, wait=False); s.synchronize(); torch.cuda.synchronize(); print("torch copy + step", round(time.perf_counter()-t0, 4))
# and the reverse: my copy + torch compute
lib = load_library()
buf = lib.pinned_array((6*S, n, n))[0]
a = torch.randn(8192, 8192, device='cuda')
def tcompute():
    for _ in range(6): b = a @ a
tcompute(); torch.cuda.synchronize()
t0 = time.perf_counter(); tcompute(); torch.cuda.synchronize(); print("torch compute alone", round(time.perf_counter()-t0, 4))
t0 = time.perf_counter(); s.download_slices3d("ex", 0, buf); s.wait_transfers(); print("my copy alone", round(time.perf_counter()-t0, 4))
t0 = time.perf_counter(); s.download_slices3d("ex", 0, buf); tcompute(); torch.cuda.synchronize(); s.wait_transfers(); print("my copy + torch compute", round(time.perf_counter()-t0, 4))
