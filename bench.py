#!/usr/bin/env python
"""bench.py -- macroparticle pushes/s and ms per xi-slice of the quasistatic wakefield sweep on B200.

Workload (BASELINE.json configs[1], SURVEY.md section 8d input 2): 3-D linear laser wake, 2048 x 512 x 512 grid,
2x2 particles per cell (P = 1 048 576 plasma macroparticles), magnetic_field_iterations = 5, a0 = 1, uniform plasma.
One bench "step" = SLICES_PER_STEP consecutive xi-slices of that sweep (push + all deposits + all field solves).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 (torchrun, one rank per GPU): the xi-loop is causally sequential, so a sweep never spans GPUs; ranks run
independent members of the a0 / density parameter scan (BASELINE.json configs[4]), no data-path collective,
"scaling": "weak".  `value` = total pushes of all ranks / max-over-ranks device time.

--impl reference times the reference's own OpenMP CPU implementation of the same path on this box's host cores:
oracle/_ref/quasistatic_pic_fast = the UNMODIFIED reference sources built by oracle/Makefile (falls back to the
plain-C oracle port, 1 thread, if that binary did not travel).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SLICES_PER_STEP = 64

WORKLOAD = dict(
    name="laser_wake_512x512_4ppc_M5",
    nx=2048, ny=512, nz=512, lx=32.0, ly=32.0, lz=32.0, ppcy=2, ppcz=2, M=5,
    a0=1.0, xsigma=4.0, ysigma=8.0, zsigma=8.0, x0=26.0,
)
# parameter scan members for N > 1 (configs[4]): (a0, rho0)
SCAN = [(1.0, -1.0), (0.5, -1.0), (2.0, -1.0), (3.0, -1.0), (1.0, -0.5), (0.5, -0.5), (2.0, -0.5), (3.0, -0.5)]


def algorithmic_bytes_per_slice(N, P, M):
    """8 * [(29 + 44 M) N + (18 + 21 M) P]"""
    return 8.0 * ((29 + 44 * M) * N + (18 + 21 * M) * P)


# algorithmic bytes of ONE launch of each kernel (DESIGN.md section "kernels"): doubles moved per cell (N) / particle (P)
KERNEL_BYTES = {
    # sort-free tile deposits (fast path); charge: slot arrays fy fz n (+ packed offset code, presence mask) -> one grid
    "k_tile_deposit_charge": lambda N, P: 8.0 * (2 * N + 3 * P),
    "k_tile_deposit_charge_mid": lambda N, P: 8.0 * (3 * N + 3.5 * P),
    "k_tile_deposit_jx": lambda N, P: 8.0 * (2 * N + 3 * P),
    "k_tile_deposit_jx_rho_chi": lambda N, P: 8.0 * (5 * N + 5 * P),
    "k_tile_deposit_rho_chi": lambda N, P: 8.0 * (3 * N + 4 * P),
    "k_tile_deposit_jy_jz": lambda N, P: 8.0 * (4 * N + 5.5 * P),  # fy fz py_next pz_next n + code; psi_middle, mask; 2 grids out
    # sorted (general) path
    "k_deposit_charge": lambda N, P: 8.0 * (2 * N + 3 * P),
    "k_deposit_jx_rho_chi": lambda N, P: 8.0 * (3 * N + 5 * P),
    "k_deposit_jx": lambda N, P: 8.0 * (2 * N + 3 * P),
    "k_deposit_rho_chi": lambda N, P: 8.0 * (3 * N + 5 * P),
    "k_deposit_jy_jz": lambda N, P: 8.0 * (2 * N + 4 * P),
    "k_particles_currents_yz": lambda N, P: 8.0 * (7.5 * P),
    # transforms
    "k_fft_rows_r2c": lambda N, P: 8.0 * (2 * N),
    "k_fft_rows_c2r": lambda N, P: 8.0 * (2 * N),
    "k_fft_cols_solve": lambda N, P: 8.0 * (3.5 * N),  # mean of the half (2N + N/2 table) and full (4N + N table) spectrum launches
    "k_potential_fixup": lambda N, P: 8.0 * (2 * N),
    "k_fft_rows_b_forward": lambda N, P: 8.0 * (9 * N),
    "k_fft_rows_b_backward": lambda N, P: 8.0 * (4 * N),
    # particles
    "k_particles_gamma": lambda N, P: 8.0 * (17.5 * P),
    "k_particles_predict": lambda N, P: 8.0 * (21.2 * P),  # predictor + corrector: reads 14.5, writes 6.5 doubles per particle (+ px in the last of the M passes)
    "k_particles_advance": lambda N, P: 8.0 * (15 * P),
    "k_particles_to_slots": lambda N, P: 8.0 * (15.5 * P),
    "k_end_of_slice": lambda N, P: 8.0 * (17 * N),
    "k_scan_chunk_sums": lambda N, P: 4.0 * N,
    "k_scan_sums": lambda N, P: 8.0 * (N / 1024),
    "k_scan_apply": lambda N, P: 4.0 * (4 * N),
    "k_sort_fill": lambda N, P: 4.0 * (3 * P),
    "k_sort_rank": lambda N, P: 4.0 * (4 * P),
    "k_next_slice": lambda N, P: 8.0,
}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the same kernels on the bench workload's transverse shape, from the
# ncu captures committed as profiles/r03_ncu_full_summary.txt (`ncu --set full`, third session: predictor, jy/jz and psi_middle
# deposits, column / row transforms, push) and profiles/r02_launches_ncu.csv (the others); cold-cache replays: an upper bound on
# the in-graph traffic
NCU_TRAFFIC = {
    "k_tile_deposit_jy_jz": 62.86e6, "k_tile_deposit_charge_mid": 32.55e6, "k_tile_deposit_jx": 26.57e6, "k_tile_deposit_charge": 26.57e6,
    "k_tile_deposit_rho_chi": 43.01e6, "k_particles_predict": 139.64e6, "k_particles_gamma": 85.90e6, "k_particles_advance": 103.83e6,
    "k_particles_to_slots": 83.92e6, "k_fft_cols_solve": 4.44e6, "k_fft_rows_b_forward": 14.70e6, "k_fft_rows_b_backward": 4.23e6,
    "k_fft_rows_r2c": 2.11e6, "k_fft_rows_c2r": 2.17e6,
}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                except ValueError:
                    continue
                for nm, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------- reference / CPU arm

REF_FAST = os.path.join(ROOT, "oracle", "_ref", "quasistatic_pic_fast")


def write_ref_toml(path, w, nx):
    dx = w["lx"] / w["nx"]
    lx = nx * dx
    with open(path, "w") as f:
        f.write(f"lx = {lx!r}\nly = {w['ly']!r}\nlz = {w['lz']!r}\ndx = {dx!r}\ndy = {w['ly'] / w['ny']!r}\ndz = {w['lz'] / w['nz']!r}\n")
        f.write(f"ppcy = {w['ppcy']}\nppcz = {w['ppcz']}\nmagnetic_field_iterations = {w['M']}\n\n[laser]\nshape = \"gaussian\"\n")
        f.write(f"a0 = {w['a0']!r}\nxsigma = {w['xsigma']!r}\nysigma = {w['ysigma']!r}\nzsigma = {w['zsigma']!r}\nx0 = {0.5 * lx!r}\n")
        f.write("\n[output]\nfields_xy = true\n")


def run_reference_sample(w, nx, threads):
    """One run of the unmodified reference on an nx-slice cut of the workload.  Returns (seconds for the nx-1 pushed
    slices, kind).  The span is first-FFT to last-FFT as recorded by the fftw3 shim (oracle/shims/fftw3_shim.cpp)."""
    P = w["ppcy"] * w["ppcz"] * w["ny"] * w["nz"]
    if os.path.exists(REF_FAST) and os.access(REF_FAST, os.X_OK):
        with tempfile.TemporaryDirectory() as wd:
            write_ref_toml(os.path.join(wd, "in.toml"), w, nx)
            env = dict(os.environ, OMP_NUM_THREADS=str(threads), QPIC_REF_TIMING_FILE=os.path.join(wd, "timing.json"))
            subprocess.run([REF_FAST, "in.toml"], cwd=wd, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            span = json.load(open(os.path.join(wd, "timing.json")))["first_fft_to_last_fft_s"]
        return span, "reference", threads, P
    # the compiled reference did not travel: time the plain-C port (single thread)
    from oracle.qo import Oracle
    from quetzalpic_b200 import gaussian_laser_widths, wake_particle_lattice
    dx = w["lx"] / w["nx"]
    o = Oracle(nx, w["ny"], w["nz"], nx * dx, w["ly"], w["lz"], w["M"])
    wx, wy, wz = gaussian_laser_widths(w["xsigma"], w["ysigma"], w["zsigma"])
    o.set_laser_gaussian(0.5 * w["a0"] ** 2, wx, wy, wz, 0.5 * nx * dx, 0.5 * w["ly"], 0.5 * w["lz"])
    o.init_wake_particles(*wake_particle_lattice(w["ny"], w["nz"], w["ly"], w["lz"], w["ppcy"], w["ppcz"]))
    o.begin()
    t0 = time.perf_counter()
    for _ in range(nx - 1):
        o.step_slice()
    span = time.perf_counter() - t0
    o.close()
    return span, "port", 1, P


def cpu_baseline(w, target_seconds=15.0):
    threads = os.cpu_count() or 1
    # calibrate with a 3-slice run, then size the sample
    span, kind, used, P = run_reference_sample(w, 3, threads)
    per_slice = span / 2
    nx = int(min(33, max(3, round(target_seconds / max(per_slice, 1e-3)) + 1)))
    if nx > 3:
        span, kind, used, P = run_reference_sample(w, nx, threads)
    pushed = nx - 1
    return {
        "value": P * pushed / span, "unit": "pushes/s", "cores": used, "kind": kind, "ms_per_slice": 1e3 * span / pushed,
        "sample": f"{pushed} xi-slices of the same 512x512 4 ppc M=5 workload ({nx}-slice box, laser centred), "
                  f"{'unmodified reference -Ofast + OpenMP, MKL FFT (DFTI in libtorch_cpu)' if kind == 'reference' else 'plain-C oracle port'}, "
                  f"OMP_NUM_THREADS={used}",
    }


def run_reference_arm(args, w):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    spans, P, kind, used = [], None, None, None
    nx = 5
    for it in range(args.warmup + args.steps):
        if it >= 1 and it < args.warmup:
            continue  # one warm-up run is enough to page the binary in; CPU runs are slow
        span, kind, used, P = run_reference_sample(w, nx, threads)
        if it >= args.warmup:
            spans.append(span)
    total = sum(spans)
    pushed = (nx - 1) * len(spans)
    value = P * pushed / total
    line = {
        "impl": "reference", "metric": "macroparticle pushes/sec (3D laser wake, per-xi-slice push+deposit+field solve)",
        "value": value, "unit": "pushes/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(spans), "ms_per_slice": 1e3 * total / pushed, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["name"], "grid": [w["nx"], w["ny"], w["nz"]], "ppc": w["ppcy"] * w["ppcz"], "particles": P,
                   "magnetic_field_iterations": w["M"], "slices_per_step": nx - 1,
                   "note": "each step = one run of the reference executable on a 5-slice cut of the workload (4 pushed slices)"},
        "cpu_baseline": {"value": value, "unit": "pushes/s", "cores": used, "kind": kind,
                         "sample": f"{len(spans)} runs x 4 xi-slices, 512x512 4 ppc M=5, OMP_NUM_THREADS={used}"},
        "e2e": {"value": value, "unit": "pushes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- B200 arm

def run_b200_arm(args, w):
    import torch
    import torch.distributed as dist
    from quetzalpic_b200 import WakefieldSolver, gaussian_laser_widths, load_library, wake_particle_lattice

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    a0, rho0 = SCAN[rank % len(SCAN)] if world > 1 else (w["a0"], -1.0)
    S, K, W = args.slices_per_step, args.steps, args.warmup
    nx, ny, nz, M = w["nx"], w["ny"], w["nz"], w["M"]
    if (K + W) * S > nx - 1:
        raise SystemExit(f"(steps + warmup) * {S} slices exceeds the {nx}-slice box")
    N = ny * nz
    wx, wy, wz = gaussian_laser_widths(w["xsigma"], w["ysigma"], w["zsigma"])
    y, z, n = wake_particle_lattice(ny, nz, w["ly"], w["lz"], w["ppcy"], w["ppcz"], rho0=rho0)
    P = y.size
    lib = load_library()

    solver = WakefieldSolver(nx, ny, nz, w["lx"], w["ly"], w["lz"], M, device=local, stop_check_interval=S)
    solver.set_laser_gaussian(0.5 * a0 * a0, wx, wy, wz, w["x0"], 0.5 * w["ly"], 0.5 * w["lz"])
    solver.init_wake_particles(y, z, n)

    # ---- device-resident throughput: inputs already in HBM, K steps timed with CUDA events on the solver's stream
    solver.begin()
    for _ in range(W):
        solver.step(S, wait=False)
    solver.synchronize()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    solver.step(K * S, wait=False)
    solver.synchronize()
    ms, launches = solver.last_step_stats()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    stopped = solver.stopped_at
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    pushes_total = float(P) * K * S * world
    value = pushes_total / (ms_max * 1e-3)
    ms_per_slice = ms_max / (K * S)

    # ---- per-kernel shares of a slice (eager, one CUDA event per launch), for the roofline of the dominant kernel
    prof = solver.profile_slices(8)
    prof_total = sum(v[1] for v in prof.values())
    kernels = []
    for name, (count, tot) in prof.items():
        per_launch_ms = tot / count
        nbytes = KERNEL_BYTES.get(name, lambda N, P: 0.0)(N, P)
        kernels.append({"kernel": name, "launches_per_slice": count / 8.0, "us_per_launch": 1e3 * per_launch_ms, "share": tot / prof_total,
                        "algorithmic_gbs": nbytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else None})
    kernels.sort(key=lambda k: -k["share"])
    peak, peak_src = measured_peak_gbs()
    top = kernels[0]
    top_bytes = KERNEL_BYTES.get(top["kernel"], lambda N, P: 0.0)(N, P)
    roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": top["algorithmic_gbs"], "peak": peak, "unit": "GB/s",
                "frac": (top["algorithmic_gbs"] or 0.0) / peak, "traffic": NCU_TRAFFIC.get(top["kernel"]), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": top_bytes, "share_of_slice": top["share"]}
    slice_bytes = algorithmic_bytes_per_slice(N, P, M)
    slice_gbs = slice_bytes / (ms_per_slice * 1e-3) / 1e9
    roofline_slice = {"bound": "hbm", "achieved": slice_gbs, "peak": peak, "unit": "GB/s", "frac": slice_gbs / peak,
                      "algorithmic_bytes_per_slice": slice_bytes, "formula": "8*[(29+44M)N+(18+21M)P] (SURVEY.md 8d)"}

    solver.close()

    # ---- end to end through the C ABI with HOST buffers.  The laser envelope is a host array as the caller holds it
    # (System_3d::a_sqr, built before the timed region like init_a_sqr in the reference; only the K*S slices this run sweeps are
    # materialised).  Two output contracts of the reference's Output_writer:
    #   "xy"  [output] fields_xy = true -- what BASELINE config #2 names: the (nx, ny) cuts of all datasets; the device evaluates the
    #         cuts (qpic_download_xy_cut), only they cross PCIe;
    #   "3d"  the six (nx, ny, nz) field arrays the reference keeps in host memory (ex, ey, ez, by, bz, susceptibility) come down
    #         every step: 805 MB per step, PCIe-bound on this box.
    T = K + W  # W untimed warm-up steps (graph instantiation, lazy module loads, first-touch of the pinned buffers), then K timed ones
    a_host, a_ptr = lib.pinned_array((T * S, ny, nz))
    xs = np.arange(nx) * (w["lx"] / nx)
    yy = np.arange(ny) * (w["ly"] / ny)
    zz = np.arange(nz) * (w["lz"] / nz)
    gy = np.exp(-((yy - 0.5 * w["ly"]) ** 2) / wy / wy)[:, None] * np.exp(-((zz - 0.5 * w["lz"]) ** 2) / wz / wz)[None, :]
    gx = 0.5 * a0 * a0 * np.exp(-((xs - w["x0"]) ** 2) / wx / wx)
    lo = nx - T * S
    for i0 in range(0, T * S, S):
        a_host[i0:i0 + S] = gx[lo + i0:lo + i0 + S, None, None] * gy[None, :, :]

    def a_block(k):
        first = nx - (k + 1) * S
        return first, a_host[first - lo:first - lo + S]

    def run_e2e(mode):
        if mode == "xy":
            fields = [("aSqr", False), ("susceptibility", False), ("ex", False), ("ey", False), ("ez", False), ("by", False), ("bz", False),
                      ("psi", True), ("rho", True), ("jx", True), ("jy", True), ("jz", True)]
            shape = (S, ny)
        else:
            fields = [(f, False) for f in ["ex", "ey", "ez", "by", "bz", "susceptibility"]]
            shape = (S, ny, nz)
        # the stop flag is polled every 8 steps here, so qpic_step only queues work and the copies of step k-1 / k+1 overlap step k
        sv = WakefieldSolver(nx, ny, nz, w["lx"], w["ly"], w["lz"], M, device=local, stop_check_interval=8 * S, keep_xy_cuts=(mode == "xy"))
        outs = [[lib.pinned_array(shape) for _ in fields] for _ in range(2)]
        probe = (0, ny // 2) if mode == "xy" else (0, ny // 2, nz // 2)
        sv.init_wake_particles(y, z, n)
        sv.begin()
        checksum = 0.0
        sv.upload_a_sqr_slices(*a_block(0))
        t0 = None
        for k in range(T):
            if k == W:
                # end of the warm-up: drain, then time exactly K steps
                sv.synchronize()
                sv.wait_transfers()
                barrier()
                t0 = time.perf_counter()
            first = nx - (k + 1) * S
            sv.step(S, wait=False)                           # block k computes ...
            if k + 1 < T:
                sv.upload_a_sqr_slices(*a_block(k + 1))      # ... while block k+1 of the envelope goes up ...
            if k >= 1 and k != W:
                sv.wait_transfers()                          # ... and block k-1's results finish landing in host memory
                checksum += float(outs[(k - 1) % 2][2][0][probe])
            for (dst, _), (f, per_slice) in zip(outs[k % 2], fields):
                if mode == "xy":
                    sv.download_xy_cut(f, first, dst, per_slice=per_slice)   # queued behind block k on the copy stream
                else:
                    sv.download_slices3d(f, first, dst)
        sv.synchronize()
        checksum += float(outs[(T - 1) % 2][2][0][probe])
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        te = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        seconds = float(te.item())
        stopped_e2e = sv.stopped_at
        for bufs in outs:
            for _, ptr in bufs:
                lib.free_pinned(ptr)
        sv.close()
        d2h = len(fields) * int(np.prod(shape)) * 8
        return {"value": pushes_total / seconds, "unit": "pushes/s", "h2d_bytes_per_step": S * N * 8, "d2h_bytes_per_step": d2h,
                "ms_per_slice": 1e3 * seconds / (K * S), "checksum": checksum, "stopped_at": stopped_e2e}

    e2e_xy = run_e2e("xy")
    e2e_3d = run_e2e("3d")
    lib.free_pinned(a_ptr)
    e2e_xy["what"] = ("output contract of the named config ([output] fields_xy = true), through the C ABI with host buffers; per step: a_sqr slab H2D "
                      "(pinned), qpic_step, the (nx, ny) Fields_xy cuts of 12 datasets D2H (pinned, cuts evaluated on the device); transfers on the "
                      "library's copy stream, double buffered, overlapping the next step's slices; W warm-up steps untimed, K steps timed")
    e2e_3d["what"] = ("same, but the six (nx, ny, nz) field arrays the reference keeps in host memory (ex, ey, ez, by, bz, susceptibility) come down "
                      "every step instead of the cuts: PCIe-bound on this box")

    if rank == 0:
        cpu = cpu_baseline(w) if (world == 1 and not args.no_cpu_baseline) else None
        line = {
            "metric": "macroparticle pushes/sec (3D laser wake, per-xi-slice push+deposit+field solve)",
            "value": value, "unit": "pushes/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K,
            "ms_per_slice": ms_per_slice, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": w["name"], "grid": [nx, ny, nz], "ppc": w["ppcy"] * w["ppcz"], "particles_per_gpu": P,
                       "magnetic_field_iterations": M, "slices_per_step": S, "a0": a0,
                       "parallelism": "1 sweep per GPU" + (f"; {world}-member a0/density scan, no collective" if world > 1 else ""),
                       "l2": "inputs larger than L2: particle SoA + per-slice rows of nine 3-D fields (>200 MB touched per slice)",
                       "stopped_at": stopped},
            "roofline": roofline, "roofline_slice": roofline_slice, "kernels": kernels[:8],
            "cpu_baseline": cpu,
            "e2e": e2e_xy, "e2e_fields_3d": e2e_3d,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--slices-per-step", type=int, default=SLICES_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="wake512", choices=["wake512", "slab4096"],
                    help="wake512 (default): BASELINE config #2 per GPU (N > 1: independent scan members); slab4096: config #4, ONE 4096^2 16 ppc "
                         "grid slab-decomposed over the N GPUs (strong scaling; tools/dist_run.py)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    w = dict(WORKLOAD)
    # keep (steps + warmup) * slices inside the 2048-slice box whatever K/W the driver asks for
    while (args.steps + args.warmup) * args.slices_per_step > w["nx"] - 1 and args.slices_per_step > 1:
        args.slices_per_step //= 2
    if args.workload == "slab4096" and args.impl == "b200":
        # config #4: one process per GPU, exchange regions mapped over CUDA IPC; a step = one xi-slice of the 268 M particle grid
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import dist_run
        sys.exit(dist_run.bench(argparse.Namespace(n=4096, nx=64, ppc=4, M=5, a0=1.0, slices=max(1, min(args.steps, 48)), warm=max(3, min(args.warmup, 8)))))
    if args.impl == "reference":
        run_reference_arm(args, w)
    else:
        run_b200_arm(args, w)


if __name__ == "__main__":
    main()
