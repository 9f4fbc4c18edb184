"""diagnostic: 2 slab ranks in one process on the visible GPUs, eager and graph mode, with timings"""
import faulthandler, os, sys, time
faulthandler.enable()
faulthandler.dump_traceback_later(100, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from cases import Case, slice_rel_l2
from quetzalpic_b200.dist import SlabGroup

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
graph = (sys.argv[2] == "graph") if len(sys.argv) > 2 else False
ndev = torch.cuda.device_count()
n = (6, 64 * (world // 2), 64)
case = Case.synthetic(n=n, d=(0.2, 0.1, 0.1), M=2, ppc=(2, 1), a0=2.0, sigma=(1.0, 1.6, 1.9))
o = case.make_oracle(); o.solve_wakefield()
t0 = time.time()
g = SlabGroup(world, n[0], n[1], n[2], case.l[0], case.l[1], case.l[2], case.M, devices=[r % ndev for r in range(world)], use_graph=graph, stop_check_interval=2)
g.set_laser_gaussian(*case.laser)
g.init_wake_particles(*case.particles)
print(f"setup {time.time()-t0:.2f}s, devices {[r % ndev for r in range(world)]}, graph={graph}", flush=True)
t0 = time.time(); g.begin(); print(f"begin {time.time()-t0:.2f}s", flush=True)
for k in range(n[0]):
    t0 = time.time()
    try:
        g.step(1)
    except Exception as e:
        print("step", k, "FAILED", e, flush=True)
        break
    print(f"step {k}: {time.time()-t0:.3f}s", flush=True)
else:
    g.each(lambda r, s: s.finish())
    for f in ["ex", "ey", "by", "susceptibility"]:
        print(f, slice_rel_l2(g.field3d(f), o.field3d(f)), flush=True)
g.close()
