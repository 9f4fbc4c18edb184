"""CPU suite (no GPU): host-side logic and the C-ABI surface.  No compute call is made without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

import quetzalpic_b200 as qp
from cases import Case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "qpic_b200.h")).read()
    declared = set(re.findall(r"\b(qpic_[a-z0-9_]+)\s*\(", header))
    declared -= {"qpic_ctx", "qpic_params", "qpic_status"}
    assert declared == set(qp.ABI_SYMBOLS), declared ^ set(qp.ABI_SYMBOLS)
    lib = qp.load_library()
    for sym in declared:
        assert hasattr(lib.L, sym), sym
    assert lib.L.qpic_abi_version() == 3


def test_library_is_cuda_only_and_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(qp.QpicError) as e:
        qp.WakefieldSolver(4, 4, 4, 1.0, 1.0, 1.0, 1)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_parameter_validation_mirrors_reference_exceptions():
    """std::invalid_argument cases of System_3d::System_3d (src/System_3d.cpp:77-85) -> QPIC_ERR_INVALID, before any CUDA call."""
    lib = qp.load_library()
    p = qp.solver.QpicParams()
    lib.L.qpic_default_params(ctypes.byref(p))
    assert (p.psi_threshold, p.magnetic_field_D, p.threshold_B, p.use_graph) == (1e-4, 10.0, 100.0, 1)
    handle = ctypes.c_void_p()
    p.nx, p.ny, p.nz, p.lx, p.ly, p.lz = 4, 4, 4, 1.0, 1.0, 1.0
    p.magnetic_field_iterations = -1
    assert lib.L.qpic_create(ctypes.byref(p), ctypes.byref(handle)) == -1
    assert b"Magnetic field iterations should be non-negative" in lib.L.qpic_last_error(None)
    p.magnetic_field_iterations = 1
    p.ny = 0
    assert lib.L.qpic_create(ctypes.byref(p), ctypes.byref(handle)) == -1
    assert lib.L.qpic_create(None, ctypes.byref(handle)) == -1


def test_wake_particle_lattice_matches_reference_init():
    """System_3d::init_wake_particles (src/System_3d.cpp:724-748): y-major order, (i+0.5)*d/ppc, n = profile/ppcy/ppcz,
    zero-density sites skipped."""
    y, z, n = qp.wake_particle_lattice(3, 2, 3.0, 1.0, 2, 1)
    assert y.size == 3 * 2 * 2
    assert np.allclose(y[:2], 0.25) and np.allclose(z[:2], [0.25, 0.75]) and np.all(n == -0.5)
    prof = qp.powerlaw2d(2, 1.0, 1.5, 0.5, min_factor=0.0)
    y, z, n = qp.wake_particle_lattice(3, 1, 3.0, 1.0, 1, 1, profile=prof)
    assert y.size == 2 and np.allclose(n, [-1.0, -1.0])  # the centre site has r = 0 -> density 0 -> skipped
    assert qp.grid_dimensions((4.8, 2.4, 20), (0.2, 0.2, 0.1)) == (23, 11, 200)


def test_oracle_and_solver_build_the_same_problem():
    case = Case.from_toml("tiny_profile")
    o = case.make_oracle()
    assert o.particle_count == case.particles[0].size
    assert np.array_equal(o.particles("n"), case.particles[2])


def _scan_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import bench
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    r, w, l = bench.dist_env()
    a0, rho0 = bench.SCAN[r % len(bench.SCAN)]
    # every rank times its own member; the job-level time is the MAX over ranks, the job-level work the SUM
    t = torch.tensor([10.0 + r], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    work = torch.tensor([1.0], dtype=torch.float64)
    dist.all_reduce(work, op=dist.ReduceOp.SUM)
    out.put((r, a0, rho0, float(t.item()), float(work.item())))
    dist.destroy_process_group()


def test_parameter_scan_sharding_world_size_2_gloo():
    """N > 1 = independent members of the a0/density scan, one per rank, no data-path collective; only the timing
    reduction (max) crosses ranks.  Exercised with gloo on CPU."""
    import torch.multiprocessing as mp
    import sys
    sys.path.insert(0, ROOT)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_scan_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    assert (res[0][1], res[0][2]) != (res[1][1], res[1][2])          # different scan members
    assert res[0][3] == res[1][3] == 11.0 and res[0][4] == res[1][4] == 2.0


def test_bench_bytes_formula():
    import bench
    # 1.554 GB per slice at config #2
    assert abs(bench.algorithmic_bytes_per_slice(512 * 512, 4 * 512 * 512, 5) - 1.554e9) < 1e6


def test_bench_per_kernel_bytes_add_up_to_the_slice_formula():
    """bench.py's per-kernel table (SURVEY 8(d) apportioned to our kernels) must add up to 8 [(29 + 44 M) N + (18 + 21 M) P] over the launches of a
    slice -- for the round-1 kernel set and for the round-2 one, where the first predictor pass does k_particles_gamma's work and k_tile_assemble
    writes the jx target (bench.py folds their shares into the kernels that took the work over)."""
    import bench
    N, P, M = 512 * 512, 4 * 512 * 512, 5
    per_slice = {"k_tile_deposit_charge": 1, "k_fft_rows_r2c": M + 1, "k_fft_cols_solve": M + 1, "k_fft_rows_c2r": M + 1, "k_particles_gamma": 1,
                 "k_tile_deposit_rho_chi": 1, "k_particles_advance": 1, "k_end_of_slice": 1, "k_particles_predict": M, "k_tile_deposit_charge_mid": M,
                 "k_tile_deposit_jy_jz": M, "k_tile_deposit_jx": M, "k_fft_rows_b_forward": M, "k_fft_cols_solve_b": M, "k_fft_rows_b_backward": M}
    total = sum(n * bench.kernel_algorithmic_bytes(k, N, P) for k, n in per_slice.items())
    assert abs(total - bench.algorithmic_bytes_per_slice(N, P, M)) < 1.0
    # the static DRAM-traffic table covers every kernel of the default slice that moves data
    default_slice = [k for k in per_slice if k not in ("k_particles_gamma", "k_tile_deposit_jx")]
    missing = [k for k in default_slice if k not in bench.NCU_DRAM_BYTES]
    assert not missing, missing
    for k, (nbytes, source) in bench.NCU_DRAM_BYTES.items():
        assert nbytes >= 0 and os.path.exists(os.path.join(os.path.dirname(bench.__file__), source)), (k, source)


# ---- host logic of the slab-decomposed sweep (quetzalpic_b200/dist.py)

def test_slab_partition_covers_every_particle_once():
    import numpy as np
    from quetzalpic_b200 import wake_particle_lattice
    from quetzalpic_b200.dist import gather_rows, half_spectrum_columns, slab_particles, slab_rows
    ny, nz, ly, lz = 64, 32, 6.4, 3.2
    y, z, n = wake_particle_lattice(ny, nz, ly, lz, 2, 2)
    for world in (2, 4):
        parts = [slab_particles(y, z, n, ny, ly, r, world) for r in range(world)]
        assert sum(p[0].size for p in parts) == y.size
        for r, (py, pz, pn) in enumerate(parts):
            r0, nr = slab_rows(ny, r, world)
            rows = (py / (ly / ny)).astype(int)
            assert rows.min() >= r0 and rows.max() < r0 + nr
        cols = [half_spectrum_columns(nz, r, world) for r in range(world)]
        assert sum(c[1] for c in cols) == nz // 2 + 1
        assert all(cols[r][0] == cols[r - 1][0] + cols[r - 1][1] for r in range(1, world) if cols[r][1] > 0)
        full = np.arange(3 * ny * nz, dtype=float).reshape(3, ny, nz)
        per_rank = []
        for r in range(world):
            a = np.full_like(full, np.nan)
            r0, nr = slab_rows(ny, r, world)
            a[:, r0:r0 + nr] = full[:, r0:r0 + nr]
            per_rank.append(a)
        assert np.array_equal(gather_rows(per_rank, ny, axis=1), full)


def _gloo_worker(rank, world, port, out):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from quetzalpic_b200 import dist as qd

        class FakeSolver:
            dist_world = world

            def dist_ipc_handle(self):
                return bytes([rank]) * 64

            def dist_attach_ipc(self, handles):
                out.put((rank, [h[0] for h in handles], [len(h) for h in handles]))

        qd.attach_process_group(FakeSolver())
    finally:
        dist.destroy_process_group()


def test_ipc_handle_exchange_over_gloo_world_size_2():
    """attach_process_group all-gathers the 64-byte handles in rank order (gloo, two CPU processes)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [(0, [0, 1], [64, 64]), (1, [0, 1], [64, 64])]


def test_beam_lattice_and_units_match_reference(golden_dir):
    """Host-side set-up of the beam path: bunch_particle_lattice restates System_3d::init_bunch_particles (reference
    src/System_3d.cpp:767-803) -- positions, momenta and order are bit-identical to the unmodified reference's Bunch_000 output, weights
    up to numpy's exp -- and plasma_units restates src/Plasma_units.h (checked against the banner the reference printed)."""
    import os
    import re
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from cases import Case
    from quetzalpic_b200 import plasma_units
    g = np.load(os.path.join(golden_dir, "tiny_qed.npz"))
    case = Case.from_toml("tiny_qed")
    p = case.bunch_particles()
    for f in ["x", "y", "z", "px", "py", "pz", "gamma"]:
        assert np.array_equal(p[f], g[f"it0/bunch/{f}"]), f
    units = plasma_units(case.base_frequency)
    weight_norm = np.prod([case.l[a] / case.n[a] for a in range(3)]) * units["number_density_norm"]
    assert np.allclose(p["n"] * weight_norm, g["it0/bunch/weight"], rtol=1e-14, atol=0)
    # three bunches, one after the other: electrons (parabolic, 2x1x2 per cell), positrons (Gaussian, every cell), photons (cylinder)
    sid = p["species_id"]
    assert list(sid[np.r_[True, sid[1:] != sid[:-1]]]) == [0, 1, 3]
    assert (sid == 1).sum() == case.n[0] * case.n[1] * case.n[2]
    banner = str(g["stdout"])
    for key, label, fmt in [("frequency", "Frequency", "{:.4g} rad/s"), ("density", "Density", "{:.4g} cm^-3"), ("wavelength", "Wavelength", "{:.4g} cm")]:
        m = re.search(label + r" (\S+) ", banner)
        assert m and float(m.group(1)) == float(fmt.split(" ")[0].format(units[key])), (key, m.group(1) if m else None)
    m = re.search(r"Schwinger field in plasma units: (\S+)", banner)
    assert float(m.group(1)) == float("{:.4g}".format(units["field_schwinger"]))
