"""DEVELOPMENT TOOL: in-graph timeline of a xi-slice from %globaltimer stamps (the -DQPIC_TIMELINE build).

ncu serialises kernels and cold-starts every launch; CUDA events between launches break the graph.  This tool replays the
captured slice graph of a timeline build (tools/libqpic_b200_timeline.so, same sources + stamps in QPIC_KERNEL_BEGIN) and
prints, per kernel: launches per slice, mean duration (dependencies satisfied -> last block out), mean "staged" time
(block 0 resident -> dependencies satisfied: how much launch latency PDL hides) and the mean idle gap between the end of
the previous kernel of the slice's critical chain and this one's start.

Usage (GPU box): python tools/timeline.py [--n 512] [--ppc 2] [--M 5] [--slices 8] [--out gpurun_out/timeline.txt]
"""
import argparse
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quetzalpic_b200.solver import QpicLibrary, WakefieldSolver, gaussian_laser_widths, wake_particle_lattice  # noqa: E402

TL_LIB = os.path.join(ROOT, "tools", "libqpic_b200_timeline.so")


def build():
    src = os.path.join(ROOT, "quetzalpic_b200", "csrc", "qpic_ctx.cu")
    if not os.path.exists(TL_LIB) or os.path.getmtime(TL_LIB) < max(
            os.path.getmtime(os.path.join(ROOT, "quetzalpic_b200", "csrc", f)) for f in os.listdir(os.path.join(ROOT, "quetzalpic_b200", "csrc"))):
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-DQPIC_TIMELINE", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
                               "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o", TL_LIB, src])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=512)
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--ppc", type=int, default=2)
    ap.add_argument("--M", type=int, default=5)
    ap.add_argument("--a0", type=float, default=1.0)
    ap.add_argument("--slices", type=int, default=8)
    ap.add_argument("--warm", type=int, default=96)
    ap.add_argument("--out", default=None)
    ap.add_argument("--dump", action="store_true", help="print every record of the first timed slice")
    ap.add_argument("--xy", action="store_true", help="keep_xy_cuts: the Fields_xy output contract of the end-to-end arm")
    a = ap.parse_args()
    build()
    lib = QpicLibrary(TL_LIB)
    L = lib.L
    L.qpic_timeline_reset.argtypes = [ctypes.c_void_p]
    L.qpic_timeline_read.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    L.qpic_timeline_read.restype = ctypes.c_long
    L.qpic_timeline_kernel_name.argtypes = [ctypes.c_int]
    L.qpic_timeline_kernel_name.restype = ctypes.c_char_p
    n, nx = a.n, a.nx
    l = 32.0
    dx = 0.015625
    s = WakefieldSolver(nx, n, n, nx * dx, l, l, a.M, stop_check_interval=10 ** 6, library=lib, keep_xy_cuts=a.xy)
    s.set_laser_gaussian(0.5 * a.a0 ** 2, *gaussian_laser_widths(4.0, 8.0, 8.0), nx * dx - 2.0, 0.5 * l, 0.5 * l)
    y, z, w = wake_particle_lattice(n, n, l, l, a.ppc, a.ppc)
    s.init_wake_particles(y, z, w)
    s.begin()
    s.step(a.warm, wait=True)
    assert L.qpic_timeline_reset(s.ctx) == 0
    s.step(a.slices, wait=True)
    ms, launches = s.last_step_stats()
    cap = 1 << 16
    buf = np.zeros(4 * cap, dtype=np.int64)
    cnt = L.qpic_timeline_read(s.ctx, buf.ctypes.data, cap)
    rec = buf[:4 * cnt].reshape(cnt, 4)
    names = {}
    starts, ends = {}, {}
    for kid, kind, t0, t1 in rec:
        names.setdefault(kid, L.qpic_timeline_kernel_name(int(kid)).decode())
        (starts if kind == 0 else ends).setdefault(kid, []).append((t0, t1))
    lines = [f"timeline: {n}x{n} grid, {a.ppc * a.ppc} ppc, M={a.M}, {a.slices} graph-replayed slices, {ms / a.slices * 1e3:.1f} us/slice "
             f"(CUDA events), {launches // a.slices} kernels/slice, PDL={"on" if os.environ.get("QPIC_PDL") == "1" else "off"}"]
    events = []  # (start_go, end, staged_at, name)
    for kid in starts:
        st = sorted(starts[kid])
        en = sorted(ends.get(kid, []))
        for i, (t0, t1) in enumerate(st):
            e = en[i][0] if i < len(en) else t1
            events.append((t1, e, t0, names[kid]))
    events.sort()
    t_first = events[0][2]
    span = max(e[1] for e in events) - t_first
    lines.append(f"span of the records: {span / a.slices / 1e3:.1f} us/slice")
    agg = {}
    for go, end, staged, name in events:
        d = agg.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += (end - go) / 1e3
        d[2] += (go - staged) / 1e3
    lines.append(f"{'kernel':34s} {'n/slice':>8s} {'dur us':>8s} {'staged us':>10s} {'us/slice':>9s}")
    tot = 0.0
    for name, (c, dur, stg) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{name:34s} {c / a.slices:8.2f} {dur / c:8.2f} {stg / c:10.2f} {dur / a.slices:9.1f}")
        tot += dur / a.slices
    lines.append(f"sum of kernel durations: {tot:.1f} us/slice (branches overlap, so this can exceed the slice time)")
    # busy / idle: union of [go, end] intervals
    iv = sorted((go, end) for go, end, _, _ in events)
    busy, cur_s, cur_e = 0, iv[0][0], iv[0][1]
    for s0, e0 in iv[1:]:
        if s0 > cur_e:
            busy += cur_e - cur_s
            cur_s, cur_e = s0, e0
        else:
            cur_e = max(cur_e, e0)
    busy += cur_e - cur_s
    lines.append(f"device busy (union of kernel intervals): {busy / a.slices / 1e3:.1f} us/slice; idle between kernels: {(span - busy) / a.slices / 1e3:.1f} us/slice")
    if a.dump:
        per = len(events) // a.slices
        lines.append("first timed slice, in order of start (us since the slice's first record): go, end, dur, staged-before-go, name")
        base = events[0][2]
        for go, end, staged, name in events[:per]:
            lines.append(f"  {(go - base) / 1e3:9.2f} {(end - base) / 1e3:9.2f} {(end - go) / 1e3:7.2f} {(go - staged) / 1e3:7.2f}  {name}")
    text = "\n".join(lines)
    print(text)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            f.write(text + "\n")
    s.close()


if __name__ == "__main__":
    main()
