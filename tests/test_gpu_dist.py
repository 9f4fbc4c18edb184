"""Slab-decomposed multi-GPU sweep (BASELINE config #4) through the C ABI, against the CPU oracle and the
single-GPU sweep.  The ranks are contexts of ONE process here (quetzalpic_b200.dist.SlabGroup); on a box with fewer GPUs than
ranks they share a device -- the exchange path (peer stores into the neighbours' regions, flag barriers, halo reduce / fill,
particle migration, distributed transposes) is the same code either way.  tools/dist_run.py is the one-process-per-GPU
(CUDA IPC + torch.distributed) version of the same check.
"""
import numpy as np
import pytest

from cases import Case, slice_rel_l2

pytestmark = pytest.mark.gpu

FIELDS = ["susceptibility", "ex", "ey", "ez", "by", "bz"]


def devices_for(world):
    import torch
    n = torch.cuda.device_count()
    return [r % n for r in range(world)]


@pytest.mark.parametrize("world,n,ppc,a0", [(2, (12, 64, 64), (2, 1), 2.0), (4, (8, 128, 64), (1, 2), 1.5), (2, (6, 128, 256), (2, 2), 1.2)])
def test_slab_sweep_matches_oracle_and_single_gpu(world, n, ppc, a0):
    from quetzalpic_b200.dist import SlabGroup
    nx, ny, nz = n
    case = Case.synthetic(n=n, d=(0.2, 0.1, 0.1), M=3, ppc=ppc, a0=a0, sigma=(1.2, 0.25 * ny * 0.1, 0.3 * nz * 0.1))
    # the laser sits a few cells off a slab edge so that particles migrate (and the ranks' counts change)
    wx, wy, wz = case.laser[1:4]
    case.laser = (0.5 * a0 * a0, wx, wy, wz, 0.6 * nx * 0.2, 0.1 * ny / world - 0.63, 0.05 * nz)
    o = case.make_oracle()
    o.solve_wakefield()
    with case.make_solver() as single:
        single.solve_wakefield()
        ref = {f: single.field3d(f) for f in FIELDS}
    y, z, w = case.particles
    with SlabGroup(world, nx, ny, nz, case.l[0], case.l[1], case.l[2], case.M, devices=devices_for(world), stop_check_interval=4) as g:
        g.set_laser_gaussian(*case.laser)
        g.init_wake_particles(y, z, w)
        before = [s.particle_count for s in g.ranks]
        stopped = g.solve_wakefield()
        after = [s.particle_count for s in g.ranks]
        assert stopped == [-1] * world
        assert sum(after) == sum(before) == y.size
        assert after != before, "the case is meant to exercise particle migration"
        for f in FIELDS:
            got = g.field3d(f)
            assert slice_rel_l2(got, o.field3d(f)) < 1e-11, f
            assert slice_rel_l2(got, ref[f]) < 1e-11, f


def test_slab_context_rejects_unsupported_shapes():
    from quetzalpic_b200 import QpicError, WakefieldSolver
    with pytest.raises(QpicError):
        WakefieldSolver(4, 100, 64, 1.0, 1.0, 1.0, 2, dist_rank=0, dist_world=2)   # not a power of two
    with pytest.raises(QpicError):
        WakefieldSolver(4, 64, 64, 1.0, 1.0, 1.0, 2, dist_rank=0, dist_world=4)    # 16 rows per rank < 32
    with pytest.raises(QpicError):
        WakefieldSolver(4, 64, 64, 1.0, 1.0, 1.0, 2, dist_rank=2, dist_world=2)    # rank out of range
    with WakefieldSolver(4, 64, 64, 1.0, 1.0, 1.0, 2, dist_rank=1, dist_world=2) as s:
        assert s.row_range == (32, 32)
        with pytest.raises(QpicError):
            s.begin()  # peers not attached


def test_stop_condition_reaches_every_rank():
    """|B| > threshold_B on one slab (src/System_3d.cpp:545-550): the stop tag travels with the next exchange barrier, every rank ends
    the sweep at the same slice and the fields written so far match the single-domain oracle."""
    from quetzalpic_b200.dist import SlabGroup
    nx, ny, nz = 10, 64, 64
    case = Case.synthetic(n=(nx, ny, nz), d=(0.2, 0.1, 0.1), M=2, ppc=(1, 1), a0=0.0)
    x = (np.arange(nx) * 0.2)[:, None, None]
    y = (np.arange(ny) * 0.1)[None, :, None]
    z = (np.arange(nz) * 0.1)[None, None, :]
    # a thin, dense driver in rank 0's slab (rows < 32): B blows up in slice 6 before the plasma has moved
    bunch = -20000.0 * np.exp(-((x - 1.2) / 0.05) ** 2 - ((y - 2.9) / 0.5) ** 2 - ((z - 3.2) / 0.5) ** 2)
    o = case.make_oracle()
    o.set_bunch_sources(bunch, bunch)
    o.solve_wakefield()
    assert o.stopped_at == 6
    with SlabGroup(2, nx, ny, nz, case.l[0], case.l[1], case.l[2], case.M, devices=devices_for(2), stop_check_interval=3) as g:
        for s in g.ranks:
            s.upload_bunch_sources(bunch, bunch)
        g.init_wake_particles(*case.particles)
        assert g.solve_wakefield() == [6, 6]
        for f in FIELDS:
            assert slice_rel_l2(g.field3d(f), o.field3d(f)) < 1e-9, f


@pytest.mark.parametrize("predict_deposit", ["0", "1"], ids=["tile-kernels", "predictor-side-deposit"])
@pytest.mark.parametrize("world", [2, 4])
def test_slab_sweep_with_half_steps_beyond_one_cell(world, predict_deposit, monkeypatch):
    """Strong drive on a coarse xi step: half steps cross more than one cell, also across slab edges (reference: any displacement is
    legal, src/array_utils.cpp:85-118).  Round 1 stopped a distributed sweep with an error here; now the far movers are listed under their
    target cell by the predictor (both kernel pairs: the tile kernels and the predictor-side deposit that config #4 selects), the psi_middle
    source deposit merges them in slot order and the 4-row halo reduce carries what lands beyond the slab edge to its owner."""
    from quetzalpic_b200.dist import SlabGroup
    monkeypatch.setenv("QPIC_PREDICT_DEPOSIT", predict_deposit)
    nx, ny, nz = 8, 64 * world // 2 * 2, 64
    case = Case.synthetic(n=(nx, ny, nz), d=(0.5, 0.1, 0.1), M=2, ppc=(2, 1), a0=3.0, sigma=(0.8, 3.0, 3.0))
    wx, wy, wz = case.laser[1:4]
    case.laser = (case.laser[0], wx, wy, wz, 0.6 * nx * 0.5, 0.1 * ny / world + 0.4, 3.2)  # next to the edge between ranks 0 and 1
    # the case is meant to have far movers (without the lists -- QPIC_FAR_MOVERS=0 -- the sweep ends with QPIC_ERR_UNSUPPORTED on slabs): in the
    # oracle's sweep some full steps exceed two cells, i.e. half steps exceed one
    probe = case.make_oracle()
    probe.begin()
    worst = 0.0
    for _ in range(nx - 1):
        y0 = probe.particles("y").copy()
        probe.step_slice()
        dy = np.abs(probe.particles("y") - y0)
        worst = max(worst, np.minimum(dy, case.l[1] - dy).max() / 0.1)
    probe.close()
    assert worst > 2.0, worst
    o = case.make_oracle()
    o.solve_wakefield()
    y, z, w = case.particles
    with SlabGroup(world, nx, ny, nz, case.l[0], case.l[1], case.l[2], case.M, devices=devices_for(world), stop_check_interval=2) as g:
        g.set_laser_gaussian(*case.laser)
        g.init_wake_particles(y, z, w)
        assert g.solve_wakefield() == [o.stopped_at] * world
        assert sum(s.particle_count for s in g.ranks) == y.size
        for f in FIELDS:
            assert slice_rel_l2(g.field3d(f), o.field3d(f)) < 1e-10, f


@pytest.mark.parametrize("dense", ["0", "1"], ids=["shuffle-sums", "shared-memory-sums"])
@pytest.mark.parametrize("world", [2, 4])
def test_slab_sweep_is_bit_reproducible(world, dense, monkeypatch):
    """Particles that migrate between ranks are appended on the receiving rank in the order the senders' atomics resolved; the cell sort
    orders them by (side, slot on the sending rank) -- sort_key in qpic_kernels.cuh -- so every per-cell sum, and with it every field of
    the sweep, is the same bit for bit from run to run (round 1: only up to the order of the particles inside a cell)."""
    from quetzalpic_b200.dist import SlabGroup
    monkeypatch.setenv("QPIC_DENSE_CELLS", dense)   # (what config #4's 16 particles per cell select)
    monkeypatch.setenv("QPIC_FOLD_RHO_CHI", dense)
    nx, ny, nz = 10, 64 * world // 2 * 2, 64
    a0 = 2.5
    case = Case.synthetic(n=(nx, ny, nz), d=(0.3, 0.1, 0.1), M=2, ppc=(2, 2), a0=a0, sigma=(1.0, 2.0, 2.5))
    wx, wy, wz = case.laser[1:4]
    case.laser = (case.laser[0], wx, wy, wz, 0.6 * nx * 0.3, 0.1 * ny / world + 0.3, 3.0)  # next to the edge between ranks 0 and 1
    y, z, w = case.particles
    runs = []
    for _ in range(3):
        with SlabGroup(world, nx, ny, nz, case.l[0], case.l[1], case.l[2], case.M, devices=devices_for(world), stop_check_interval=3) as g:
            g.set_laser_gaussian(*case.laser)
            g.init_wake_particles(y, z, w)
            before = [s.particle_count for s in g.ranks]
            assert g.solve_wakefield() == [-1] * world
            assert [s.particle_count for s in g.ranks] != before, "the case is meant to exercise particle migration"
            runs.append({f: g.field3d(f) for f in FIELDS})
    for f in FIELDS:
        assert np.array_equal(runs[0][f], runs[1][f]) and np.array_equal(runs[0][f], runs[2][f]), f
