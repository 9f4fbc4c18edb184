"""Shared helpers for the parity tests: build the same problem for the CPU oracle and the CUDA solver."""
import os
import re

import numpy as np

from quetzalpic_b200 import (SPECIES, bunch_particle_lattice, bunch_profile_from_table, concatenate_bunches, gaussian_laser_widths,
                             grid_dimensions, plasma_units, powerlaw2d, wake_particle_lattice)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def parse_toml(path):
    """Tiny reader for the reference-style input files used as fixtures (key = value, [table])."""
    root, cur = {}, None
    cur = root
    for line in open(path):
        line = line.split("#", 1)[0].strip()
        if not line:
            continue
        m = re.match(r"\[\[(\w+)\]\]", line)
        if m:  # array of tables
            cur = {}
            root.setdefault(m.group(1), []).append(cur)
            continue
        m = re.match(r"\[(\w+)\]", line)
        if m:
            cur = root.setdefault(m.group(1), {})
            continue
        k, v = [s.strip() for s in line.split("=", 1)]
        if v in ("true", "false"):
            cur[k] = v == "true"
        elif v[0] in "\"'":
            cur[k] = v[1:-1]
        else:
            cur[k] = float(v) if any(c in v for c in ".eE") else int(v)
    return root


class Case:
    """Everything System_3d derives from an input file before solve_wakefield (reference src/System_3d.cpp:64-197,
    src/Config_reader.cpp:18-262)."""

    def __init__(self, cfg):
        self.cfg = cfg
        l = (float(cfg["lx"]), float(cfg["ly"]), float(cfg["lz"]))
        d = (float(cfg["dx"]), float(cfg["dy"]), float(cfg["dz"]))
        self.l = l
        self.n = grid_dimensions(l, d)
        self.M = int(cfg.get("magnetic_field_iterations", 0))
        self.ppcy = int(cfg.get("ppcy", 1))
        self.ppcz = int(cfg.get("ppcz", 1))
        self.laser = None
        if "laser" in cfg:
            ls = cfg["laser"]
            wx, wy, wz = gaussian_laser_widths(ls["xsigma"], ls["ysigma"], ls["zsigma"])
            self.laser = (0.5 * ls["a0"] * ls["a0"], wx, wy, wz, float(ls.get("x0", ls["xsigma"])),
                          float(ls.get("y0", 0.5 * l[1])), float(ls.get("z0", 0.5 * l[2])))
        profile, rho0 = None, -1.0
        if "plasma" in cfg:
            pl = cfg["plasma"]
            if pl["profile"] == "powerlaw":
                profile = powerlaw2d(pl["power"], pl["radius"], pl.get("y0", 0.5 * l[1]), pl.get("z0", 0.5 * l[2]),
                                     pl.get("min_factor", 0.0), pl.get("max_factor", 1e10), -1.0)
            else:
                rho0 = float(pl.get("rho0", -1))
        nx, ny, nz = self.n
        self.particles = wake_particle_lattice(ny, nz, l[1], l[2], self.ppcy, self.ppcz, profile=profile, rho0=rho0)
        self.photons = []  # (photons, max radiation probability) per step of a qed time loop

    # -- beam side of System_3d (reference src/System_3d.cpp:106-162, :199-236)
    def bunch_particles(self):
        """bunch_particles of System_3d for the [bunch] / [[bunch]] tables of the input"""
        tables = self.cfg.get("bunch", [])
        tables = tables if isinstance(tables, list) else [tables]
        bunches = []
        for t in tables:
            ppc = (int(t.get("ppcx", 1)), int(t.get("ppcy", 1)), int(t.get("ppcz", 1)))
            bunches.append(bunch_particle_lattice(self.n, self.l, ppc, bunch_profile_from_table(t, self.l), float(t["gamma"]),
                                                  t.get("species", "electron")))
        return concatenate_bunches(bunches)

    @property
    def base_frequency(self):
        """reference src/Config_reader.cpp:29-39, src/Plasma_units.h:11-13"""
        if "base_density_CGS" in self.cfg:
            pi, e, m = 3.141592653589793238, 4.803204672997e-10, 9.1093837015e-28
            return float(np.sqrt(4 * pi * e * e * float(self.cfg["base_density_CGS"]) / m))
        return float(self.cfg.get("base_frequency_SI", 1.0))

    @property
    def time_iterations(self):
        dt = float(self.cfg.get("dt", self.l[0]))
        return int(float(self.cfg.get("t_end", 0.0)) / dt) + 1, dt

    def reference_uniforms(self, species_ids, steps):
        """The rand_double() values (reference src/monte_carlo.cpp:7-9) a ONE-thread run of the reference consumes in
        advance_bunch_particles with qed = true (src/System_3d.cpp:890, :897): two per charged particle, in particle order, step
        after step, from glibc's rand() in its default state (the reference never calls srand).  Returns [(u_emit, u_accept)] per step."""
        import ctypes
        libc = ctypes.CDLL(None)
        libc.rand.restype = ctypes.c_int
        libc.srand(1)  # == the state of a process that never seeded
        rand_max = 2147483647
        charged = np.array([not s[2] for s in SPECIES.values()])[np.asarray(species_ids)]
        out = []
        for _ in range(steps):
            u1, u2 = np.zeros(len(charged)), np.zeros(len(charged))
            for i in np.nonzero(charged)[0]:
                u1[i] = libc.rand() / rand_max
                u2[i] = libc.rand() / rand_max
            out.append((u1, u2))
        return out

    def run_time_loop(self, make_backend, on_iteration):
        """System_3d::run (reference src/System_3d.cpp:199-236) on either backend.  make_backend() -> (solver, beam) where
        solver has init_wake_particles / solve_wakefield and beam has deposit_bunch_densities / deposit_wakefield_to_bunch /
        advance_bunch_particles; on_iteration(ti, solver, beam) is called where the reference writes Bunch_NNN."""
        iterations, dt = self.time_iterations
        units = plasma_units(self.base_frequency)
        solver, beam = make_backend()
        for ti in range(iterations):
            beam.deposit_bunch_densities()
            solver.init_wake_particles(*self.particles)  # solve_wakefield re-creates the plasma lattice, :262
            solver.solve_wakefield()
            beam.deposit_wakefield_to_bunch(units["field_schwinger"])
            on_iteration(ti, solver, beam)
            if ti < iterations - 1:
                if self.cfg.get("qed", False):
                    if ti == 0:
                        uniforms = self.reference_uniforms(self.bunch_particles()["species_id"], iterations - 1)
                    self.photons.append(beam.advance_bunch_particles_qed(dt, units["frequency"], *uniforms[ti]))
                else:
                    beam.advance_bunch_particles(dt)
        return solver, beam

    @classmethod
    def from_toml(cls, name):
        return cls(parse_toml(os.path.join(GOLDEN, "inputs", name + ".toml")))

    @classmethod
    def synthetic(cls, n, d, M, ppc=(1, 1), a0=1.0, sigma=None, x0=None, **extra):
        nx, ny, nz = n
        cfg = dict(lx=nx * d[0], ly=ny * d[1], lz=nz * d[2], dx=d[0], dy=d[1], dz=d[2], ppcy=ppc[0], ppcz=ppc[1],
                   magnetic_field_iterations=M)
        if a0:
            sigma = sigma or (0.5 * cfg["lx"], 0.5 * cfg["ly"], 0.5 * cfg["lz"])
            cfg["laser"] = dict(a0=a0, xsigma=sigma[0], ysigma=sigma[1], zsigma=sigma[2], x0=x0 if x0 is not None else 0.6 * cfg["lx"])
        cfg.update(extra)
        c = cls(cfg)
        c.n = (nx, ny, nz)  # exact by construction
        return c

    def make_oracle(self, **kw):
        from oracle.qo import Oracle
        nx, ny, nz = self.n
        o = Oracle(nx, ny, nz, self.l[0], self.l[1], self.l[2], self.M, **kw)
        if self.laser:
            o.set_laser_gaussian(*self.laser)
        o.init_wake_particles(*self.particles)
        return o

    def make_solver(self, a_sqr=None, **kw):
        from quetzalpic_b200 import WakefieldSolver
        nx, ny, nz = self.n
        s = WakefieldSolver(nx, ny, nz, self.l[0], self.l[1], self.l[2], self.M, **kw)
        if a_sqr is not None:
            s.upload_a_sqr(a_sqr)
        elif self.laser:
            s.set_laser_gaussian(*self.laser)
        s.init_wake_particles(*self.particles)
        return s


def rel_l2(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def slice_rel_l2(a, b):
    """max over slices of ||a_i - b_i|| / max_i ||b_i|| (early slices can be ~1e-90: SURVEY.md section 8c)."""
    diff = np.sqrt(((a - b) ** 2).sum(axis=(1, 2)))
    ref = np.sqrt((b ** 2).sum(axis=(1, 2))).max()
    return diff.max() / (ref if ref > 0 else 1.0)
