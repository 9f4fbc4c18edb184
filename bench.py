#!/usr/bin/env python
"""bench.py -- macroparticle pushes/s and ms per xi-slice of the quasistatic wakefield sweep on B200.

Workload (BASELINE.json configs[1], SURVEY.md section 8d input 2): 3-D linear laser wake, 2048 x 512 x 512 grid,
2x2 particles per cell (P = 1 048 576 plasma macroparticles), magnetic_field_iterations = 5, a0 = 1, uniform plasma.
One bench "step" = SLICES_PER_STEP consecutive xi-slices of that sweep (push + all deposits + all field solves).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 (torchrun, one rank per GPU): the xi-loop is causally sequential, so a sweep never spans GPUs; ranks run
independent members of the a0 / density parameter scan (BASELINE.json configs[4]), no data-path collective,
"scaling": "weak".  `value` = total pushes of all ranks / max-over-ranks device time.  The same JSON line then carries, for N > 1,
`scan_members` (per-rank ms, a0, rho0, fast-path aborts), `control_identical_members` (all ranks on the SAME member: separates interference
from member cost) and `slab4096`: BASELINE config #4 (64 x 4096^2, 16 ppc) slab-decomposed over the N GPUs (tools/dist_run.py: the
real multi-GPU path -- distributed transposes, halo exchange, particle migration over NVLink peer memory), with a parity probe against
the single-GPU kernels.

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


# SURVEY.md section 8(d) apportioned to OUR kernels: (doubles per cell, doubles per particle) of ONE launch.  The split follows the
# derivation printed under the formula there (DESIGN.md section 6 has the table); by construction
#     sum over the launches of a slice = 8 * [(29 + 44 M) N + (18 + 21 M) P]
# (checked at run time: `roofline.kernels_sum_check`).  A kernel that does NOT appear in the 8(d) derivation (cell sort, slot
# copies) has algorithmic bytes 0: its traffic is overhead by 8(d)'s own statement ("sort / reorder traffic ... not counted").
SURVEY_BYTES = {
    # once per slice: psi solve 12 = deposit target 1 + rows 5 + columns 2 + rows 2 + normalise/clamp 2 (fused into the backward rows)
    "k_tile_deposit_charge": (1, 3), "k_fft_rows_r2c": (5, 0), "k_fft_cols_solve": (2, 0), "k_fft_rows_c2r": (4, 0),
    "k_particles_gamma": (2, 7),            # gathers 2; reads y z py pz n, writes gamma px
    "k_tile_deposit_rho_chi": (5, 0), "k_tile_fold_rho_chi": (5, 0),       # rho / jx / chi targets of the once-per-slice deposit (its particle reads are in the 7 above)
    "k_particles_advance": (1, 8), "k_end_of_slice": (9, 0),
    # per B iteration: predictor 10 + corrector share 6 of the "11" (the other 5 = the jy/jz deposit's reads); gathers 4 + 2
    "k_particles_predict": (6, 16),
    "k_tile_deposit_charge_mid": (1, 0),    # deposit target of the psi_middle solve (particle reads are in the predictor's 10)
    "k_tile_deposit_jy_jz": (3, 5), "k_tile_deposit_jx": (5, 0),
    "k_fft_rows_b_forward": (10, 0), "k_fft_cols_solve_b": (4, 0), "k_fft_rows_b_backward": (4, 0),
}

# dram__bytes_read.sum + dram__bytes_write.sum per launch on the bench workload's transverse shape (512 x 512, P = 2^20), from the
# `ncu --set full` captures committed under profiles/ -- STATIC numbers from those files (cold-cache serialised replays: an upper
# bound on the in-graph traffic), not measured by this run.  {kernel: (bytes, file)}
NCU_DRAM_BYTES = {}
_ncu_table = os.path.join(ROOT, "profiles", "ncu_dram_bytes.json")
if os.path.exists(_ncu_table):
    try:
        NCU_DRAM_BYTES = {k: (float(v["bytes"]), v["source"]) for k, v in json.load(open(_ncu_table)).items()}
    except Exception:
        NCU_DRAM_BYTES = {}


def kernel_algorithmic_bytes(name, N, P):
    n, p = SURVEY_BYTES.get(name, (0, 0))
    return 8.0 * (n * N + p * P)


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

def slab_record(args, rank, world, local):
    """BASELINE config #4 (64 x 4096^2, 16 ppc, M = 5) slab-decomposed over the `world` GPUs of this launch (strong scaling):
    driver-timed numbers for the real multi-GPU data path (peer-store transposes, halo reduce / fill, particle migration,
    k_xchg_barrier), plus a parity probe of the same exchange code on a 256-row analogue against a single-GPU sweep."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dist_run
    from cases import Case, slice_rel_l2
    from quetzalpic_b200 import gaussian_laser_widths
    from quetzalpic_b200.dist import slab_rows

    out = {}
    # ---- parity probe: 4 x 256 x 256, 2 x 2 ppc, M = 3, laser on a slab edge (particles migrate); 3 pushed slices
    n = (4, 256, 256)
    case = Case.synthetic(n=n, d=(0.2, 0.1, 0.1), M=3, ppc=(2, 2), a0=1.8, sigma=(1.2, 6.0, 6.0))
    wx, wy, wz = case.laser[1:4]
    r0e, _ = slab_rows(n[1], min(1, world - 1), world)
    laser = (case.laser[0], wx, wy, wz, 0.5, 0.1 * r0e + 0.05, 12.8)
    s = dist_run.make_rank_solver(n[0], n[1], n[2], case.l[0], case.l[1], case.l[2], case.M, laser, (2, 2), rank, world, local, stop_check_interval=4)
    s.solve_wakefield()
    fields = ["ex", "ey", "ez", "by", "bz", "susceptibility"]
    r0, nr = slab_rows(n[1], rank, world)
    mine = {f: s.field3d(f)[:, r0:r0 + nr].copy() for f in fields}
    live = int(s.particle_count)
    s.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, live))
    if rank == 0:
        case.laser = laser
        with case.make_solver(device=local) as single:
            single.solve_wakefield()
            worst = max(slice_rel_l2(np.concatenate([g[0][f] for g in gathered], axis=1), single.field3d(f)) for f in fields)
        out["parity_probe"] = {"grid": list(n), "ppc": 4, "M": 3, "max_field_error_vs_single_gpu": worst, "tolerance": 1e-11,
                               "particles_per_rank_after": [g[1] for g in gathered], "particles_total": int(case.particles[0].size)}
        if not (worst < 1e-11 and sum(g[1] for g in gathered) == case.particles[0].size):
            raise SystemExit(f"bench.py: slab parity probe failed: {out['parity_probe']}")
    dist.barrier()

    # ---- throughput: config #4
    nx, ng, M, ppc = 64, 4096, 5, 4
    d = 64.0 / 4096
    lx, l = nx * d, ng * d
    wx, wy, wz = gaussian_laser_widths(0.5, 0.25 * l, 0.25 * l)
    laser = (0.5, wx, wy, wz, 0.6 * lx, 0.5 * l, 0.5 * l)
    t0 = time.time()
    s = dist_run.make_rank_solver(nx, ng, ng, lx, l, l, M, laser, (ppc, ppc), rank, world, local, stop_check_interval=10 ** 6)
    setup_s = time.time() - t0
    P_total = float(ng) * ng * ppc * ppc
    warm, slices = 3, args.slab_slices
    s.begin()
    s.step(warm, wait=False)
    s.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    s.step(slices, wait=False)
    s.synchronize()
    ms, launches = s.last_step_stats()
    torch.cuda.synchronize()
    dist.barrier()
    per_rank = [None] * world
    dist.all_gather_object(per_rank, (float(ms), int(s.particle_count), int(s.fast_path_aborts), int(s.stopped_at)))
    s.close()
    if rank == 0:
        ms_max = max(r[0] for r in per_rank)
        N = float(ng) * ng
        alg = algorithmic_bytes_per_slice(N, P_total, M)
        peak, _ = measured_peak_gbs()
        gbs = alg / (ms_max / slices * 1e-3) / 1e9
        out.update({
            "workload": "laser_wake_4096x4096_16ppc_M5 (BASELINE config #4), y-slabs, peer stores / loads over NVLink inside the slice graph, no NCCL on the data path",
            "grid": [nx, ng, ng], "particles": P_total, "scaling": "strong", "n_gpus": world, "slices": slices, "warmup_slices": warm,
            "ms_per_slice": ms_max / slices, "value": P_total * slices / (ms_max * 1e-3), "unit": "pushes/s",
            "ms_per_rank": [r[0] / slices for r in per_rank], "particles_per_rank_after": [r[1] for r in per_rank],
            "fast_path_aborts": [r[2] for r in per_rank], "stopped_at": [r[3] for r in per_rank],
            "achieved_gbs_aggregate": gbs, "achieved_gbs_per_gpu": gbs / world, "roofline_frac_per_gpu": gbs / world / peak,
            "kernels_per_slice": launches // slices, "setup_s": setup_s,
        })
    return out


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

    def gather(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    a0, rho0 = SCAN[rank % len(SCAN)] if world > 1 else (w["a0"], -1.0)
    S, K, W = args.slices_per_step, args.steps, args.warmup
    nx, ny, nz, M = w["nx"], w["ny"], w["nz"], w["M"]
    if (K + W) * S > nx - 1:
        raise SystemExit(f"(steps + warmup) * {S} slices exceeds the {nx}-slice box")
    T = K + W
    N = ny * nz
    wx, wy, wz = gaussian_laser_widths(w["xsigma"], w["ysigma"], w["zsigma"])
    lib = load_library()
    P = ny * nz * w["ppcy"] * w["ppcz"]
    # slices whose results the end-to-end arm probes (first slice of every block except the last warm-up one, see run_e2e)
    probe_slices = [nx - (k + 1) * S for k in range(T) if k != W - 1]

    def device_resident(a0_, rho0_, want_profile):
        """inputs already in HBM, K steps timed with CUDA events on the solver's stream; max over ranks"""
        y, z, n = wake_particle_lattice(ny, nz, w["ly"], w["lz"], w["ppcy"], w["ppcz"], rho0=rho0_)
        solver = WakefieldSolver(nx, ny, nz, w["lx"], w["ly"], w["lz"], M, device=local, stop_check_interval=S)
        solver.set_laser_gaussian(0.5 * a0_ * a0_, wx, wy, wz, w["x0"], 0.5 * w["ly"], 0.5 * w["lz"])
        solver.init_wake_particles(y, z, n)
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
        res = {"ms": float(ms), "launches": int(launches), "stopped_at": int(solver.stopped_at), "fast_path_aborts": int(solver.fast_path_aborts),
               "a0": a0_, "rho0": rho0_, "clocks": clocks}
        # what the end-to-end arm must reproduce: ex (its Fields_xy cut at z = lz/2 falls on plane nz/2 exactly) and ez on the probe slices
        res["probe_ex"] = [float(solver.slice3d("ex", i)[ny // 2, nz // 2]) for i in probe_slices]
        res["probe_ez"] = [float(solver.slice3d("ez", i)[ny // 2, nz // 2]) for i in probe_slices]
        if want_profile:
            res["profile"] = solver.profile_slices(8)  # eager slices, one CUDA event per launch
        solver.close()
        return res

    main_run = device_resident(a0, rho0, True)
    per_rank = gather({k: main_run[k] for k in ("ms", "a0", "rho0", "fast_path_aborts", "stopped_at")})
    ms_max = max(r["ms"] for r in per_rank)
    pushes_total = float(P) * K * S * world
    value = pushes_total / (ms_max * 1e-3)
    ms_per_slice = ms_max / (K * S)
    launches, stopped, clocks = main_run["launches"], main_run["stopped_at"], main_run["clocks"]
    control = None
    if world > 1:
        # control: every rank runs the SAME member (a0 = 1, rho0 = -1), so that the curve separates interference between the
        # ranks (shared host, power) from the cost differences between scan members
        ctl_run = device_resident(w["a0"], -1.0, False)
        ctl_rank = gather(ctl_run["ms"])
        control = {"what": f"{world} identical members (a0 = {w['a0']}, rho0 = -1): same timed region", "ms_per_slice_per_rank": [m / (K * S) for m in ctl_rank],
                   "ms_per_slice": max(ctl_rank) / (K * S), "value": pushes_total / (max(ctl_rank) * 1e-3), "unit": "pushes/s"}

    # ---- roofline: SURVEY.md 8(d) slice formula over the graph-timed slice; per-kernel table from the eager profile
    peak, peak_src = measured_peak_gbs()
    prof = main_run["profile"]
    prof_total = sum(v[1] for v in prof.values())
    kernels, alg_sum, dram_sum, dram_known = [], 0.0, 0.0, True
    for name, (count, tot) in prof.items():
        per_launch_ms = tot / count
        nbytes = kernel_algorithmic_bytes(name, N, P)
        if name == "k_particles_predict" and "k_particles_gamma" not in prof and M > 0:
            # the first predictor pass of a slice does k_particles_gamma's work (predict_inputs<true>): its 8(d) share is spread over the M launches
            nbytes += kernel_algorithmic_bytes("k_particles_gamma", N, P) / M
        if name == "k_tile_deposit_charge_mid" and "k_tile_deposit_jx" not in prof:
            # predictor-side deposit with the jx deposit riding along: k_tile_assemble<true> writes the jx target as well
            nbytes += kernel_algorithmic_bytes("k_tile_deposit_jx", N, P)
        if name == "k_tile_deposit_rho_chi" and "k_tile_fold_rho_chi" in prof:
            nbytes = 0.0  # launched, but returns at once unless the |B| stop ended the slice early: k_tile_fold_rho_chi wrote the targets
        dram = NCU_DRAM_BYTES.get(name)
        alg_sum += nbytes * count / 8.0
        if dram:
            dram_sum += dram[0] * count / 8.0
        elif per_launch_ms * 1e3 > 4.0:
            dram_known = False
        kernels.append({"kernel": name, "launches_per_slice": count / 8.0, "us_per_launch_eager": 1e3 * per_launch_ms, "share": tot / prof_total,
                        "algorithmic_bytes": nbytes, "algorithmic_gbs": nbytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else None,
                        "frac_of_peak": nbytes / (per_launch_ms * 1e-3) / 1e9 / peak if per_launch_ms > 0 else None,
                        "ncu_dram_bytes": dram[0] if dram else None, "ncu_dram_bytes_source": (dram[1] + " (static_from_profile)") if dram else None})
    kernels.sort(key=lambda k: -k["share"])
    slice_bytes = algorithmic_bytes_per_slice(N, P, M)
    slice_gbs = slice_bytes / (ms_per_slice * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": slice_gbs, "peak": peak, "unit": "GB/s", "frac": slice_gbs / peak,
                "traffic": dram_sum if (dram_known and dram_sum > 0) else None,
                "traffic_note": "sum over the slice's launches of ncu dram__bytes_read+write per launch, static_from_profile (profiles/ncu_dram_bytes.json), cold-cache replays",
                "what": "one launch = one replay of the slice graph (all kernels of one xi-slice); achieved = algorithmic bytes per slice / graph-timed ms per slice",
                "algorithmic_bytes_per_slice": slice_bytes, "formula": "8*[(29+44M)N+(18+21M)P] (SURVEY.md 8d)", "peak_source": peak_src,
                "kernels_sum_check": alg_sum / slice_bytes, "eager_us_per_slice": 1e3 * prof_total / 8.0, "kernels": kernels[:12]}

    # ---- end to end through the C ABI with HOST buffers.  The laser envelope is a host array as the caller holds it
    # (System_3d::a_sqr, built before the timed region like init_a_sqr in the reference; only the K*S slices this run sweeps are
    # materialised).  Two output contracts of the reference's Output_writer:
    #   "xy"  [output] fields_xy = true -- what BASELINE config #2 names: the (nx, ny) cuts of all datasets; the device evaluates the
    #         cuts (qpic_download_xy_cut), only they cross PCIe;
    #   "3d"  the six (nx, ny, nz) field arrays the reference keeps in host memory (ex, ey, ez, by, bz, susceptibility) come down
    #         every step: 805 MB per step, PCIe-bound on this box.
    y, z, n = wake_particle_lattice(ny, nz, w["ly"], w["lz"], w["ppcy"], w["ppcz"], rho0=rho0)
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
        # the abort / stop flags are polled once per step (as in the device-resident arm); qpic_step still only queues work, the copies
        # of step k-1 / k+1 overlap step k
        sv = WakefieldSolver(nx, ny, nz, w["lx"], w["ly"], w["lz"], M, device=local, stop_check_interval=args.e2e_poll_steps * S, keep_xy_cuts=(mode == "xy"))
        outs = [[lib.pinned_array(shape) for _ in fields] for _ in range(2)]
        probe = (0, ny // 2) if mode == "xy" else (0, ny // 2, nz // 2)
        sv.init_wake_particles(y, z, n)
        sv.begin()
        probes = []
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
                probes.append(float(outs[(k - 1) % 2][2][0][probe]))
            for (dst, _), (f, per_slice) in zip(outs[k % 2], fields):
                if mode == "xy":
                    sv.download_xy_cut(f, first, dst, per_slice=per_slice)   # queued behind block k on the copy stream
                else:
                    sv.download_slices3d(f, first, dst)
        sv.synchronize()
        probes.append(float(outs[(T - 1) % 2][2][0][probe]))
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        seconds_rank = gather(t1 - t0)
        seconds = max(seconds_rank)
        stopped_e2e, aborts = sv.stopped_at, int(sv.fast_path_aborts)
        for bufs in outs:
            for _, ptr in bufs:
                lib.free_pinned(ptr)
        sv.close()
        # the device-resident arm swept the same slices of the same problem: the host-buffer path must land on the same numbers
        # (a_sqr comes from numpy's exp here and from the device's there: agreement to rounding, not bit for bit)
        expect = main_run["probe_ex" if mode == "xy" else "probe_ez"]
        scale = max(max(abs(v) for v in expect), 1e-300)
        err = max(abs(a - b) for a, b in zip(probes, expect)) / scale
        if stopped_e2e < 0 and stopped < 0 and not (err < 1e-9):
            raise SystemExit(f"bench.py: e2e ({mode}) results differ from the device-resident sweep of the same slices: rel. error {err:.3e}\n"
                             f"  e2e   {probes}\n  resident {expect}")
        d2h = len(fields) * int(np.prod(shape)) * 8
        return {"value": pushes_total / seconds, "unit": "pushes/s", "h2d_bytes_per_step": S * N * 8, "d2h_bytes_per_step": d2h,
                "ms_per_slice": 1e3 * seconds / (K * S), "checksum": float(sum(probes)), "checksum_expected": float(sum(expect)),
                "checksum_rel_error": err, "stopped_at": stopped_e2e, "ms_per_slice_per_rank": [1e3 * t / (K * S) for t in seconds_rank],
                "fast_path_aborts_per_rank": gather(aborts)}

    e2e_xy = run_e2e("xy")
    e2e_3d = run_e2e("3d")
    lib.free_pinned(a_ptr)
    e2e_xy["what"] = ("output contract of the named config ([output] fields_xy = true), through the C ABI with host buffers; per step: a_sqr slab H2D "
                      "(pinned), qpic_step, the (nx, ny) Fields_xy cuts of 12 datasets D2H (pinned, cuts evaluated on the device); transfers on the "
                      "library's copy stream, double buffered, overlapping the next step's slices; W warm-up steps untimed, K steps timed")
    e2e_3d["what"] = ("same, but the six (nx, ny, nz) field arrays the reference keeps in host memory (ex, ey, ez, by, bz, susceptibility) come down "
                      "every step instead of the cuts: PCIe-bound on this box")

    slab = None
    if world > 1 and not args.no_slab:
        slab = slab_record(args, rank, world, local)

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
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e_xy, "e2e_fields_3d": e2e_3d,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world > 1:
            line["scan_members"] = [{"rank": r, "a0": m["a0"], "rho0": m["rho0"], "ms_per_slice": m["ms"] / (K * S), "fast_path_aborts": m["fast_path_aborts"],
                                     "stopped_at": m["stopped_at"]} for r, m in enumerate(per_rank)]
            line["control_identical_members"] = control
            line["slab4096"] = slab
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--slices-per-step", type=int, default=SLICES_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-slab", action="store_true", help="N > 1: skip the slab-decomposed config #4 record")
    ap.add_argument("--slab-slices", type=int, default=16, help="timed xi-slices of the slab-decomposed config #4 record (N > 1)")
    ap.add_argument("--e2e-poll-steps", type=int, default=1, help="end-to-end arm: poll the device's stop / abort flags every this many steps")
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
