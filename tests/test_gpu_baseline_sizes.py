"""Parity at the BASELINE.json problem sizes (through the C ABI): the launch shapes the big grids use -- 8192-block particle kernels,
1024+-block tile kernels, the power-of-two transform engine at 512 / 1024 / 4096 points, the predictor-side deposit that switches
on at >= 2^20 cells, the slab-decomposed transposes over 4096-point axes -- against the CPU oracle (pinned bit for bit to the
unmodified reference at this size, tests/test_oracle_cpu.py) and against a golden summary of the reference's own output.

Tolerances as in test_gpu_parity.py: 1e-12 per slice from identical state, 1e-9 over a sweep (north_star: 1e-12 / 1e-6).
"""
import os

import numpy as np
import pytest

from cases import Case, rel_l2, slice_rel_l2
from test_oracle_cpu import summary_errors

pytestmark = pytest.mark.gpu

FIELDS = ["susceptibility", "ex", "ey", "ez", "by", "bz"]
SLICE_FIELDS = ["psi", "rho", "jx", "jy", "jz"]


def sweep_errors(s, o):
    worst = {}
    for f in FIELDS:
        worst[f] = slice_rel_l2(s.field3d(f), o.field3d(f))
    for f in SLICE_FIELDS:
        worst[f] = slice_rel_l2(s.field3d(f, per_slice=True), o.field3d(f, per_slice=True))
    for f in ["y", "z", "py", "pz"]:
        worst["p." + f] = rel_l2(s.particles(f), o.particles(f))
    return worst


def test_config2_transverse_problem_matches_oracle_and_reference(golden_dir):
    """BASELINE config #2 (512 x 512, 2 x 2 ppc = 1 048 576 particles, M = 5, a0 = 1) on a 6-slice box through the laser:
    (a) every field of every slice and the final particle state against the oracle, 1e-10 (a sweep; the lock-step test below holds 1e-12);
    (b) the same run against the summary of the UNMODIFIED reference's output (axis, transverse lines, decimated planes, norms)."""
    g = np.load(os.path.join(golden_dir, "baseline512_6slices_summary.npz"))
    case = Case.from_toml("baseline512_6slices")
    assert case.n == (6, 512, 512)
    o = case.make_oracle()
    o.solve_wakefield()
    with case.make_solver(keep_slices_3d=True, keep_xy_cuts=True) as s:
        assert s.solve_wakefield() == o.stopped_at == -1
        assert s.fast_path_aborts == 0
        worst = sweep_errors(s, o)
        # a 5-slice sweep (particle state NOT re-synchronised between slices): sweep tolerance; by / bz are ~1e-5 of the other fields here
        # (the wake has barely started), measured 1.2e-12 on them and <= 6e-13 on everything else
        bad = {k: v for k, v in worst.items() if not v <= 1e-10}
        assert not bad, f"beyond 1e-10: {bad} (all: {worst})"
        for f in ["aSqr"] + FIELDS + SLICE_FIELDS:
            a = s.field3d(f, per_slice=f in SLICE_FIELDS)
            assert summary_errors(a, g, f, 2, 1) < 1e-11, f
            ref_xy = g[f + "/xy"]
            got_xy = s.xy_cut(f, per_slice=f in SLICE_FIELDS)[:, ::8]
            scale = max(abs(float(g[f + "/min"])), abs(float(g[f + "/max"])), 1e-300)  # (the cut of ez lies in its symmetry plane: noise)
            assert np.abs(got_xy - ref_xy).max() <= 1e-11 * scale, f


def test_config2_per_slice_parity_from_identical_state():
    """Lock-step at 512 x 512 x 4 ppc, M = 5: the oracle's state is injected before every slice, both sides advance one slice,
    everything the slice produced agrees to 1e-12.  Pre-heated plasma (seeded random momenta, jittered lattice) so that cells hold
    0..9 particles and ~10 % of the particles change cell in a half step: the mover lists of the tile deposits and the ragged
    slot ranges are exercised at the full launch shapes."""
    rng = np.random.default_rng(512)
    case = Case.synthetic(n=(4, 512, 512), d=(0.015625, 0.0625, 0.0625), M=5, ppc=(2, 2), a0=1.0, sigma=(4.0, 8.0, 8.0), x0=0.03)
    y, z, n = case.particles
    y = np.mod(y + rng.uniform(-0.02, 0.02, y.size), case.l[1])
    z = np.mod(z + rng.uniform(-0.02, 0.02, z.size), case.l[2])
    case.particles = (y, z, n)
    nx = case.n[0]
    o = case.make_oracle()
    o.begin()
    py0, pz0 = rng.normal(0.0, 0.4, y.size), rng.normal(0.0, 0.4, y.size)
    with case.make_solver(a_sqr=o.field3d("aSqr"), keep_slices_3d=True) as s:
        s.begin()
        for i in range(nx - 1, 0, -1):
            if i == nx - 1:
                o.set_particles("py", py0)
                o.set_particles("pz", pz0)
            for f in ["y", "z", "py", "pz"]:
                s.set_particles(f, o.particles(f))
            s.set_field2d("psi_prev", o.field2d("psi"))
            for f in ["jy", "jz"]:
                s.set_field2d(f, o.field2d(f))
            by_prev = o.field3d("by")[i + 1] if i < nx - 1 else np.zeros(case.n[1:])
            bz_prev = o.field3d("bz")[i + 1] if i < nx - 1 else np.zeros(case.n[1:])
            s.set_field2d("by_cur", by_prev)
            s.set_field2d("bz_cur", bz_prev)
            s.current_slice = i
            assert o.step_slice() == 0
            s.step(1)
            # (the first slice of a neutral plasma has psi, ex ~ 1e-17 per node: rounding noise on both sides, so errors are measured
            # against at least 1e-3 per node -- the fields of this case are O(1e-2 .. 1))
            floor = 1e-3 * np.sqrt(case.n[1] * case.n[2])
            errs = {}
            for f in FIELDS:
                ref = o.field3d(f)[i]
                errs[f] = np.linalg.norm(s.slice3d(f, i) - ref) / max(np.linalg.norm(ref), floor)
            for f in SLICE_FIELDS:
                ref = o.field3d(f, per_slice=True)[i]
                errs[f + "_s"] = np.linalg.norm(s.slice3d(f, i, per_slice=True) - ref) / max(np.linalg.norm(ref), floor)
            for f in ["y", "z", "py", "pz"]:
                errs["p." + f] = rel_l2(s.particles(f), o.particles(f))
            assert max(errs.values()) <= 1e-12, f"slice {i}: {errs}"
        assert s.fast_path_aborts == 0


def test_config3_transverse_problem_1024_blowout_drive():
    """BASELINE config #3's transverse problem (1024 x 1024, 2 x 2 ppc = 4 194 304 particles, M = 10, a0 = 4) on a 3-slice box.
    >= 2^20 cells: the predictor-side psi_middle deposit (k_particles_predict_dep + k_tile_assemble) is chosen automatically, the
    column transforms run 1024 points.  Jittered lattice (ragged cells).  Sweep tolerance 1e-9, in practice ~1e-14."""
    rng = np.random.default_rng(1024)
    case = Case.synthetic(n=(3, 1024, 1024), d=(0.015625, 0.03125, 0.03125), M=10, ppc=(2, 2), a0=4.0, sigma=(4.0, 8.0, 8.0), x0=0.02)
    y, z, n = case.particles
    y = np.mod(y + rng.uniform(-0.01, 0.01, y.size), case.l[1])
    z = np.mod(z + rng.uniform(-0.01, 0.01, z.size), case.l[2])
    case.particles = (y, z, n)
    o = case.make_oracle()
    o.solve_wakefield()
    with case.make_solver(keep_slices_3d=True) as s:
        assert s.solve_wakefield() == o.stopped_at == -1
        worst = sweep_errors(s, o)
        bad = {k: v for k, v in worst.items() if not v <= 1e-9}
        assert not bad, f"beyond 1e-9: {bad} (all: {worst})"
        assert max(worst.values()) < 1e-11, worst


@pytest.mark.parametrize("world,n", [(2, (3, 128, 4096)), (4, (3, 128, 4096)), (2, (3, 4096, 128))])
def test_slab_sweep_with_4096_point_axes(world, n):
    """BASELINE config #4's axis length on the slab-decomposed path: 4096-point row transforms whose spectrum is dealt out over the
    ranks' column buffers (2049 half-spectrum columns over 2 / 4 ranks), and 4096-point distributed column transforms."""
    import torch
    from quetzalpic_b200.dist import SlabGroup
    nx, ny, nz = n
    devs = [r % torch.cuda.device_count() for r in range(world)]
    case = Case.synthetic(n=n, d=(0.2, 0.1, 0.1), M=3, ppc=(1, 1), a0=1.6, sigma=(1.2, 0.2 * ny * 0.1, 0.2 * nz * 0.1))
    wx, wy, wz = case.laser[1:4]
    case.laser = (case.laser[0], wx, wy, wz, 0.4, 0.1 * ny / world - 0.33, 0.05 * nz)
    o = case.make_oracle()
    o.solve_wakefield()
    y, z, w = case.particles
    with SlabGroup(world, nx, ny, nz, case.l[0], case.l[1], case.l[2], case.M, devices=devs, stop_check_interval=4) as g:
        g.set_laser_gaussian(*case.laser)
        g.init_wake_particles(y, z, w)
        assert g.solve_wakefield() == [-1] * world
        assert sum(s.particle_count for s in g.ranks) == y.size
        for f in FIELDS:
            assert slice_rel_l2(g.field3d(f), o.field3d(f)) < 1e-11, f


def test_config4_density_selects_the_dense_cell_kernels_single_gpu():
    """BASELINE config #4's density (4 x 4 = 16 particles per cell) on a 256 x 256 grid, 4-slice box through the laser: at >= 8 particles per
    cell the library selects, by itself, the shared-memory form of the predictor's per-cell sums and rho / susceptibility from the last
    predictor pass (k_particles_predict_dep<., 2, true> + k_tile_fold_rho_chi) -- nothing is forced through the environment here."""
    case = Case.synthetic(n=(4, 256, 256), d=(0.015625, 0.125, 0.125), M=5, ppc=(4, 4), a0=1.5, sigma=(4.0, 8.0, 8.0), x0=0.03)
    o = case.make_oracle()
    o.solve_wakefield()
    with case.make_solver(keep_slices_3d=True) as s:
        assert s.solve_wakefield() == -1
        worst = sweep_errors(s, o)
    assert max(worst.values()) < 1e-11, worst


def test_config4_density_on_slabs():
    """... and on two slabs (the rule uses the density of the particles a rank was handed): 16 particles per cell, migration across the slab edge."""
    import torch
    from quetzalpic_b200.dist import SlabGroup
    nx, ny, nz = 4, 256, 128
    case = Case.synthetic(n=(nx, ny, nz), d=(0.2, 0.1, 0.1), M=3, ppc=(4, 4), a0=1.8, sigma=(1.2, 5.0, 4.0))
    wx, wy, wz = case.laser[1:4]
    case.laser = (case.laser[0], wx, wy, wz, 0.4, 0.1 * ny / 2 - 0.33, 0.05 * nz)
    o = case.make_oracle()
    o.solve_wakefield()
    y, z, w = case.particles
    devs = [r % torch.cuda.device_count() for r in range(2)]
    with SlabGroup(2, nx, ny, nz, case.l[0], case.l[1], case.l[2], case.M, devices=devs, stop_check_interval=4) as g:
        g.set_laser_gaussian(*case.laser)
        g.init_wake_particles(y, z, w)
        assert g.solve_wakefield() == [-1, -1]
        assert sum(s.particle_count for s in g.ranks) == y.size
        for f in FIELDS:
            assert slice_rel_l2(g.field3d(f), o.field3d(f)) < 1e-11, f
