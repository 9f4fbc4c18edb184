"""Parity of the beam ("bunch") path on the GPU -- deposit_bunch_densities, deposit_wakefield_to_bunch, advance_bunch_particles
and the time loop of System_3d::run (reference src/System_3d.cpp:199-236, :843-886, :926-946) -- through the C ABI, against the
golden fixture of the unmodified reference and the CPU oracle (oracle/qo_bunch.c).  FP64 tolerance: 1e-12 relative L2 for one
step from identical state, 1e-9 over the three-iteration time loop (the sweeps in between are the laminar wake solver)."""
import os

import numpy as np
import pytest

from cases import Case, rel_l2, slice_rel_l2

pytestmark = pytest.mark.gpu

BUNCH_OUT = ["x", "y", "z", "px", "py", "pz", "gamma", "ex", "ey", "ez", "by", "bz", "chi"]


def test_beam_time_loop_matches_reference_fixture(golden_dir):
    """tests/golden/tiny_beam.npz: Fields_NNN and Bunch_NNN of three time iterations of the unmodified reference (electron
    driver with ppc 2x1x2, positron bunch filling the box, photon beam)"""
    from quetzalpic_b200 import plasma_units
    g = np.load(os.path.join(golden_dir, "tiny_beam.npz"))
    case = Case.from_toml("tiny_beam")
    particles = case.bunch_particles()
    weight_norm = np.prod([case.l[a] / case.n[a] for a in range(3)]) * plasma_units(1.0)["number_density_norm"]
    worst = {}

    def make_backend():
        s = case.make_solver(keep_slices_3d=True)
        s.bunch_init(particles)
        return s, s

    def check(ti, s, beam):
        for f in ["rho_bunch", "jx_bunch", "ex", "ey", "ez", "by", "bz", "susceptibility", "aSqr"]:
            worst[(ti, f)] = slice_rel_l2(s.field3d(f), g[f"it{ti}/f3/{f}"])
        for f in ["psi", "rho", "jx", "jy", "jz"]:
            worst[(ti, f)] = slice_rel_l2(s.field3d(f, per_slice=True), g[f"it{ti}/f3/{f}"])
        for f in BUNCH_OUT:
            worst[(ti, "bunch." + f)] = rel_l2(s.bunch(f), g[f"it{ti}/bunch/{f}"])
        worst[(ti, "bunch.weight")] = rel_l2(s.bunch("n") * weight_norm, g[f"it{ti}/bunch/weight"])

    s, _ = case.run_time_loop(make_backend, check)
    s.close()
    bad = {k: v for k, v in worst.items() if not v <= 1e-9}
    assert not bad, bad
    assert len({k[0] for k in worst}) == 3


def random_beam(rng, n, l, count, spread=1.0, weight=1.0):
    """beam particles all over the box, including outside it: x beyond both ends (dropped by deposit and gather), y and z up
    to one box length outside (periodic images)"""
    x = rng.uniform(-0.1 * l[0], 1.1 * l[0], count)
    y = rng.uniform(-spread * l[1], (1 + spread) * l[1], count)
    z = rng.uniform(-spread * l[2], (1 + spread) * l[2], count)
    # particles exactly on nodes and on the last x plane
    x[:4] = [0.0, l[0] / n[0] * (n[0] - 1), l[0] / n[0] * (n[0] - 2), l[0]]
    y[:4] = [0.0, l[1] / n[1] * 3, l[1], -l[1] / n[1]]
    px = rng.uniform(5.0, 50.0, count)
    py, pz = rng.normal(0, 1.0, count), rng.normal(0, 1.0, count)
    sid = rng.integers(0, 4, count).astype(np.int32)
    gamma = np.sqrt(np.where(sid == 3, 0.0, 1.0) + px * px + py * py + pz * pz)  # physical states: chi is real
    return dict(x=x, y=y, z=z, px=px, py=py, pz=pz, gamma=gamma, n=weight * rng.uniform(0.1, 2.0, count), species_id=sid)


@pytest.mark.parametrize("n,count", [((12, 10, 14), 5000), ((9, 16, 32), 20000), ((5, 3, 7), 300), ((2, 6, 6), 100)])
def test_beam_steps_match_oracle_from_identical_state(n, count):
    """each of the three beam steps against the oracle from the same state; fields to gather from are random arrays"""
    from oracle.qo import Oracle, OracleBunch
    from quetzalpic_b200 import SPECIES, WakefieldSolver
    rng = np.random.default_rng(count)
    l = (0.3 * n[0], 0.25 * n[1], 0.2 * n[2])
    particles = random_beam(rng, n, l, count, weight=0.02 * n[0] * n[1] * n[2] / count)  # a weak driver: the sweep stays laminar
    o = Oracle(*n, *l, 2)
    ob = OracleBunch(o, particles, list(SPECIES.values()))
    ob.deposit_bunch_densities()
    with WakefieldSolver(*n, *l, 2) as s:
        s.bunch_init(particles)
        s.deposit_bunch_densities()
        for f in ["rho_bunch", "jx_bunch"]:
            got, ref = s.field3d(f), o.field3d(f)
            assert rel_l2(got, ref) < 1e-12, f
        # total deposited charge = sum over the particles inside [0, (nx-1) dx) of charge * n (every weight set sums to 1)
        charge = np.array([v[0] for v in SPECIES.values()])[particles["species_id"]]
        inside = (np.floor(particles["x"] / (l[0] / n[0])) >= 0) & (np.floor(particles["x"] / (l[0] / n[0])) + 1 < n[0])
        assert abs(s.field3d("rho_bunch").sum() - (charge * particles["n"])[inside].sum()) < 1e-9 * count
        # bit-reproducible: no floating-point atomics
        first = s.field3d("jx_bunch")
        s.deposit_bunch_densities()
        assert np.array_equal(first, s.field3d("jx_bunch"))
        # gather: the solver's field arrays are filled by a sweep, so run one on both sides with the beam as the driver
        lattice = Case.synthetic(n=n, d=tuple(l[a] / n[a] for a in range(3)), M=2, a0=0).particles
        o.init_wake_particles(*lattice)
        o.solve_wakefield()
        s.init_wake_particles(*lattice)
        s.solve_wakefield()
        ob.deposit_wakefield_to_bunch(3.0)
        s.deposit_wakefield_to_bunch(3.0)
        for f in ["ex", "ey", "ez", "by", "bz", "chi"]:
            assert rel_l2(s.bunch(f), ob[f]) < 1e-11, f
        # push from identical fields at the particles
        for f in ["ex", "ey", "ez", "by", "bz"]:
            s.set_bunch(f, ob[f])
        ob.advance_bunch_particles(0.7)
        s.advance_bunch_particles(0.7)
        for f in ["x", "y", "z", "px", "py", "pz", "gamma"]:
            assert rel_l2(s.bunch(f), ob[f]) < 1e-14, f


def test_beam_deposit_large_grid_properties():
    """512 x 256 x 256 cells, 4 M beam particles: size-independent properties -- charge conservation, determinism, linearity in
    the weights"""
    from quetzalpic_b200 import SPECIES, WakefieldSolver
    n, l = (512, 256, 256), (8.0, 16.0, 16.0)
    rng = np.random.default_rng(7)
    count = 1 << 22
    particles = random_beam(rng, n, l, count, spread=0.0)
    particles["x"] = np.clip(rng.normal(4.0, 0.6, count), 0.0, l[0] - 2 * l[0] / n[0])
    particles["y"] = np.clip(rng.normal(8.0, 0.5, count), 0.0, l[1] * (1 - 1e-9))  # dense core: hundreds of particles per cell
    particles["z"] = np.clip(rng.normal(8.0, 0.5, count), 0.0, l[2] * (1 - 1e-9))
    charge = np.array([v[0] for v in SPECIES.values()])[particles["species_id"]]
    with WakefieldSolver(*n, *l, 0) as s:
        s.bunch_init(particles)
        s.deposit_bunch_densities()
        rho = s.field3d("rho_bunch")
        assert abs(rho.sum() - (charge * particles["n"]).sum()) < 1e-9 * np.abs(charge * particles["n"]).sum()
        s.deposit_bunch_densities()
        assert np.array_equal(rho, s.field3d("rho_bunch"))
        s.set_bunch("n", 2.0 * particles["n"])  # exact in binary: every product doubles exactly
        s.deposit_bunch_densities()
        assert np.array_equal(2.0 * rho, s.field3d("rho_bunch"))


def test_beam_call_order_errors():
    from quetzalpic_b200 import QpicError, WakefieldSolver
    with WakefieldSolver(4, 8, 8, 1.0, 2.0, 2.0, 1) as s:
        with pytest.raises(QpicError):
            s.deposit_bunch_densities()  # before bunch_init
        bad = dict(x=[0.5], y=[0.5], z=[0.5], px=[1.0], py=[0.0], pz=[0.0], gamma=[2.0], n=[1.0], species_id=[7])
        with pytest.raises(QpicError):
            s.bunch_init(bad)  # species id outside the table
        s.init_wake_particles(*Case.synthetic(n=(4, 8, 8), d=(0.25, 0.25, 0.25), M=1, a0=0).particles)
        s.solve_wakefield()
        ok = dict(bad, species_id=[0])
        if "emu" in os.path.basename(s.lib.path):
            return  # development emulator: no CUDA graph, so late set-up is harmless there
        with pytest.raises(QpicError):
            s.bunch_init(ok)  # rho_bunch / jx_bunch cannot become resident after the slice graph was captured


def test_qed_time_loop_matches_reference_fixture(golden_dir):
    """qed = true (tests/golden/tiny_qed.npz, the unmodified reference on one thread): with the uniform numbers that run drew from
    glibc's rand(), the device emits the same photons in the same order and leaves the same beam; 1e-9 over the time loop"""
    from quetzalpic_b200 import plasma_units
    g = np.load(os.path.join(golden_dir, "tiny_qed.npz"))
    case = Case.from_toml("tiny_qed")
    particles = case.bunch_particles()
    weight_norm = np.prod([case.l[a] / case.n[a] for a in range(3)]) * plasma_units(case.base_frequency)["number_density_norm"]
    worst = {}

    def make_backend():
        s = case.make_solver()
        s.bunch_init(particles)
        return s, s

    def check(ti, s, beam):
        for f in BUNCH_OUT:
            worst[(ti, "bunch." + f)] = rel_l2(s.bunch(f), g[f"it{ti}/bunch/{f}"])

    s, _ = case.run_time_loop(make_backend, check)
    s.close()
    assert len(case.photons) == 2
    for ti, (ph, maxp) in enumerate(case.photons):
        assert len(ph["x"]) == len(g[f"it{ti}/photons/x"]) > 20
        for f in ["x", "y", "z", "px", "py", "pz", "gamma", "ex", "chi"]:
            worst[(ti, "photons." + f)] = rel_l2(ph[f], g[f"it{ti}/photons/{f}"])
        worst[(ti, "photons.weight")] = rel_l2(ph["n"] * weight_norm, g[f"it{ti}/photons/weight"])
        assert f"Max radiation probability {maxp:.3g}" in str(g["stdout"])
    bad = {k: v for k, v in worst.items() if not v <= 1e-9}
    assert not bad, bad


def test_qed_step_matches_oracle_and_device_rng_statistics():
    """one emission step from identical state against the oracle (supplied uniforms, weak and strong chi branches of W_new), then the
    device generator: reproducible per (seed, step), different across steps, and an emission count within 5 sigma of the
    expectation sum_p E_u[min(1, P_p(u))] estimated with numpy uniforms through the oracle"""
    from oracle.qo import Oracle, OracleBunch
    from quetzalpic_b200 import SPECIES, WakefieldSolver
    rng = np.random.default_rng(11)
    n, l, count = (8, 12, 12), (2.0, 3.0, 3.0), 40000
    particles = random_beam(rng, n, l, count)
    particles["px"] *= 40.0
    particles["gamma"] = np.sqrt(np.where(particles["species_id"] == 3, 0.0, 1.0) + particles["px"] ** 2 + particles["py"] ** 2 + particles["pz"] ** 2)
    fields = {f: rng.normal(0, 0.3, count) for f in ["ex", "ey", "ez", "by", "bz"]}
    chi = np.abs(rng.normal(0, 0.2, count))  # both sides of 0.13333
    chi[:50] = 0.0
    frequency, dt = 5.64e13, 0.8
    u1, u2 = rng.uniform(0, 1, count), rng.uniform(0, 1, count) * 0.3

    def fresh_oracle():
        ob = OracleBunch(Oracle(*n, *l, 0), particles, list(SPECIES.values()))
        for f, v in fields.items():
            ob[f][:] = v
        ob["chi"][:] = chi
        return ob

    ob = fresh_oracle()
    ref_photons, ref_max = ob.advance_bunch_particles_qed(dt, frequency, u1, u2)
    assert 500 < len(ref_photons["x"]) < count

    def fresh_solver():
        s = WakefieldSolver(*n, *l, 0)
        s.bunch_init(particles)
        for f, v in fields.items():
            s.set_bunch(f, v)
        s.set_bunch("chi", chi)
        return s

    with fresh_solver() as s:
        photons, maxp = s.advance_bunch_particles_qed(dt, frequency, u1, u2)
        assert len(photons["x"]) == len(ref_photons["x"])
        assert abs(maxp - ref_max) <= 1e-6 * ref_max  # single-precision powf inside W_new
        for f in ["x", "y", "z", "px", "py", "pz", "gamma", "n"]:
            assert rel_l2(photons[f], ref_photons[f]) < 1e-13, f
        for f in ["x", "y", "z", "px", "py", "pz", "gamma"]:
            assert rel_l2(s.bunch(f), ob[f]) < 1e-13, f
    # device-side generator
    counts = []
    for seed in (1, 1, 2):
        with fresh_solver() as s:
            ph, _ = s.advance_bunch_particles_qed(dt, frequency, seed=seed)
            first = ph["px"].copy()
            ph2, _ = s.advance_bunch_particles_qed(dt, frequency, seed=seed)  # next step: new numbers
            counts.append((len(first), first.sum(), len(ph2["px"])))
    assert counts[0][:2] == counts[1][:2] and counts[0][:2] != counts[2][:2]
    expect, var = 0.0, 0.0
    trials = 16
    for t in range(trials):  # Monte-Carlo estimate of the expected number of emissions with independent numpy uniforms
        ob = fresh_oracle()
        ph, _ = ob.advance_bunch_particles_qed(dt, frequency, rng.uniform(0, 1, count), rng.uniform(0, 1, count))
        expect += len(ph["x"]) / trials
        var += len(ph["x"]) ** 2 / trials
    sigma = max(np.sqrt(max(var - expect ** 2, 0.0)), np.sqrt(expect))
    assert abs(counts[0][0] - expect) < 5 * sigma + 5, (counts, expect, sigma)
