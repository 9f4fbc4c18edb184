"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the CPU parity oracle (oracle/qpic_oracle.c).

Nothing under quetzalpic_b200/ imports this module.  It is used by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs, and only as the checker.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libqpic_oracle.so")

FIELDS_2D = ["psi", "psi_prev", "psi_middle", "rho", "jx", "jy", "jz", "jy_next", "jz_next", "djy_dxi", "djz_dxi", "rho_ion"]
FIELDS_3D = ["aSqr", "susceptibility", "rho_bunch", "jx_bunch", "ex", "ey", "ez", "by", "bz",
             "psi", "rho", "jx", "jy", "jz"]  # the last five: per-slice outputs, Fields-file layout
PARTICLE_FIELDS = ["y", "z", "py", "pz", "n", "gamma", "px", "py_next", "pz_next", "y_middle", "z_middle"]


class QoParams(ctypes.Structure):
    _fields_ = [
        ("nx", ctypes.c_int), ("ny", ctypes.c_int), ("nz", ctypes.c_int),
        ("lx", ctypes.c_double), ("ly", ctypes.c_double), ("lz", ctypes.c_double),
        ("magnetic_field_iterations", ctypes.c_int),
        ("psi_threshold", ctypes.c_double),
        ("magnetic_field_D", ctypes.c_double),
        ("threshold_B", ctypes.c_double),
    ]


def build(force=False):
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        L.qo_create.restype = ctypes.c_void_p
        L.qo_create.argtypes = [ctypes.POINTER(QoParams)]
        L.qo_destroy.argtypes = [ctypes.c_void_p]
        L.qo_set_laser_gaussian.argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 7
        L.qo_set_a_sqr.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.qo_set_bunch_sources.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.qo_init_wake_particles.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        L.qo_init_wake_particles_uniform.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double]
        L.qo_solve_wakefield.argtypes = [ctypes.c_void_p]
        L.qo_begin.argtypes = [ctypes.c_void_p]
        L.qo_step_slice.argtypes = [ctypes.c_void_p]
        L.qo_finish.argtypes = [ctypes.c_void_p]
        L.qo_current_slice.argtypes = [ctypes.c_void_p]
        L.qo_stopped_at.argtypes = [ctypes.c_void_p]
        L.qo_particle_count.restype = ctypes.c_size_t
        L.qo_particle_count.argtypes = [ctypes.c_void_p]
        L.qo_particles.restype = dp
        L.qo_particles.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.qo_field2d_ptr.restype = dp
        L.qo_field2d_ptr.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.qo_field3d_ptr.restype = dp
        L.qo_field3d_ptr.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.qo_deposit.argtypes = [ctypes.c_double] * 3 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 4
        L.qo_deposit_raw.argtypes = [ctypes.c_double] * 3 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 2
        L.qo_gather.restype = ctypes.c_double
        L.qo_gather.argtypes = [ctypes.c_double] * 2 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 4
        L.qo_gather_yder.restype = ctypes.c_double
        L.qo_gather_yder.argtypes = [ctypes.c_double] * 2 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 2
        L.qo_gather_zder.restype = ctypes.c_double
        L.qo_gather_zder.argtypes = [ctypes.c_double] * 2 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 2
        L.qo_solve_poisson.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double]
        _lib = L
    return _lib


class Oracle:
    """One oracle context.  Field / particle accessors return numpy VIEWS of the oracle's memory."""

    def __init__(self, nx, ny, nz, lx, ly, lz, magnetic_field_iterations, psi_threshold=1e-4, magnetic_field_D=10.0,
                 threshold_B=100.0):
        self.L = lib()
        self.params = QoParams(nx, ny, nz, lx, ly, lz, magnetic_field_iterations, psi_threshold, magnetic_field_D, threshold_B)
        self.ctx = self.L.qo_create(ctypes.byref(self.params))
        if not self.ctx:
            raise RuntimeError("qo_create failed")
        self.nx, self.ny, self.nz = nx, ny, nz

    def close(self):
        if self.ctx:
            self.L.qo_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup
    def set_laser_gaussian(self, amplitude, wx, wy, wz, x0, y0, z0):
        self.L.qo_set_laser_gaussian(self.ctx, amplitude, wx, wy, wz, x0, y0, z0)

    def set_a_sqr(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == (self.nx, self.ny, self.nz)
        self.L.qo_set_a_sqr(self.ctx, a.ctypes.data)

    def set_bunch_sources(self, rho_bunch=None, jx_bunch=None):
        r = None if rho_bunch is None else np.ascontiguousarray(rho_bunch, dtype=np.float64)
        j = None if jx_bunch is None else np.ascontiguousarray(jx_bunch, dtype=np.float64)
        self.L.qo_set_bunch_sources(self.ctx, None if r is None else r.ctypes.data, None if j is None else j.ctypes.data)

    def init_wake_particles(self, y, z, n):
        y = np.ascontiguousarray(y, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        n = np.ascontiguousarray(n, dtype=np.float64)
        if self.L.qo_init_wake_particles(self.ctx, y.ctypes.data, z.ctypes.data, n.ctypes.data, y.size) != 0:
            raise RuntimeError("qo_init_wake_particles failed")

    def init_wake_particles_uniform(self, ppcy, ppcz, rho0=-1.0):
        if self.L.qo_init_wake_particles_uniform(self.ctx, ppcy, ppcz, rho0) != 0:
            raise RuntimeError("qo_init_wake_particles_uniform failed")

    # -- sweep
    def solve_wakefield(self):
        return self.L.qo_solve_wakefield(self.ctx)

    def begin(self):
        self.L.qo_begin(self.ctx)

    def step_slice(self):
        return self.L.qo_step_slice(self.ctx)

    def finish(self):
        self.L.qo_finish(self.ctx)

    @property
    def current_slice(self):
        return self.L.qo_current_slice(self.ctx)

    @property
    def stopped_at(self):
        return self.L.qo_stopped_at(self.ctx)

    # -- state views
    @property
    def particle_count(self):
        return self.L.qo_particle_count(self.ctx)

    def particles(self, name):
        p = self.L.qo_particles(self.ctx, PARTICLE_FIELDS.index(name))
        return np.ctypeslib.as_array(p, shape=(max(self.particle_count, 1),))[: self.particle_count]

    def set_particles(self, name, value):
        """overwrite one particle field (the views above alias the oracle's memory)"""
        self.particles(name)[:] = np.asarray(value, dtype=np.float64)

    def field2d(self, name):
        p = self.L.qo_field2d_ptr(self.ctx, FIELDS_2D.index(name))
        return np.ctypeslib.as_array(p, shape=(self.ny, self.nz))

    def field3d(self, name, per_slice=False):
        """`per_slice=True` selects the Fields-file style accumulation of the 2-D arrays psi/rho/jx/jy/jz."""
        idx = FIELDS_3D.index(name, 9) if per_slice else FIELDS_3D.index(name)
        p = self.L.qo_field3d_ptr(self.ctx, idx)
        return np.ctypeslib.as_array(p, shape=(self.nx, self.ny, self.nz))


# ---------------------------------------------------------------------------------------------- beam particles (oracle/qo_bunch.c)
BUNCH_FIELDS = ["x", "y", "z", "px", "py", "pz", "gamma", "n", "ex", "ey", "ez", "by", "bz", "chi"]


class QoBunch(ctypes.Structure):
    _fields_ = [("count", ctypes.c_size_t)] + [(k, ctypes.c_void_p) for k in BUNCH_FIELDS] + [("species_id", ctypes.c_void_p)]


class OracleBunch:
    """bunch_particles of System_3d and the three beam-side steps of System_3d::run (reference src/System_3d.cpp:199-236),
    operating on an Oracle's 3-D arrays.  species: list of (charge, mass, is_photon) by species id."""

    def __init__(self, oracle, particles, species):
        self.o = oracle
        L = self.L = lib()
        vp, ci, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        L.qo_bunch_deposit_densities.argtypes = [vp, vp, ci, ci, ci, dbl, dbl, dbl, vp, vp]
        L.qo_bunch_gather_fields.argtypes = [vp, ci, ci, ci, dbl, dbl, dbl, vp, vp, vp, vp, vp, dbl]
        L.qo_bunch_advance.argtypes = [vp, vp, vp, dbl]
        count = len(particles["species_id"])
        self.a = {k: np.zeros(count) for k in BUNCH_FIELDS}
        for k in ["x", "y", "z", "px", "py", "pz", "gamma", "n"]:
            self.a[k][:] = particles[k]
        self.species_id = np.ascontiguousarray(particles["species_id"], dtype=np.int32)
        self.charge = np.array([s[0] for s in species], dtype=np.float64)
        self.cmr = np.array([0.0 if s[1] == 0.0 else s[0] / s[1] for s in species], dtype=np.float64)
        self.photon = np.array([1 if s[2] else 0 for s in species], dtype=np.int32)
        self.s = QoBunch(count, *[self.a[k].ctypes.data for k in BUNCH_FIELDS], self.species_id.ctypes.data)
        p = oracle.params
        self.geom = (p.nx, p.ny, p.nz, p.lx / p.nx, p.ly / p.ny, p.lz / p.nz)

    def __getitem__(self, name):
        return self.a[name]

    def deposit_bunch_densities(self):
        self.L.qo_bunch_deposit_densities(ctypes.byref(self.s), self.charge.ctypes.data, *self.geom,
                                          self.o.field3d("rho_bunch").ctypes.data, self.o.field3d("jx_bunch").ctypes.data)

    def deposit_wakefield_to_bunch(self, field_schwinger):
        f = [self.o.field3d(k).ctypes.data for k in ["ex", "ey", "ez", "by", "bz"]]
        self.L.qo_bunch_gather_fields(ctypes.byref(self.s), *self.geom, *f, float(field_schwinger))

    def advance_bunch_particles(self, dt):
        self.L.qo_bunch_advance(ctypes.byref(self.s), self.cmr.ctypes.data, self.photon.ctypes.data, float(dt))

    def advance_bunch_particles_qed(self, dt, frequency, u_emit, u_accept):
        """qed = true (reference src/System_3d.cpp:887-921); returns (photons dict, max radiation probability)"""
        L = self.L
        L.qo_bunch_advance_qed.restype = ctypes.c_size_t
        L.qo_bunch_advance_qed.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_double,
                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        count = self.s.count
        ph = {k: np.zeros(max(count, 1)) for k in BUNCH_FIELDS}
        sid = np.zeros(max(count, 1), dtype=np.int32)
        ps = QoBunch(0, *[ph[k].ctypes.data for k in BUNCH_FIELDS], sid.ctypes.data)
        u1 = np.ascontiguousarray(u_emit, dtype=np.float64)
        u2 = np.ascontiguousarray(u_accept, dtype=np.float64)
        assert u1.shape == u2.shape == (count,)
        maxp = ctypes.c_double(0.0)
        n = L.qo_bunch_advance_qed(ctypes.byref(self.s), self.cmr.ctypes.data, self.photon.ctypes.data, float(dt), float(frequency),
                                   u1.ctypes.data, u2.ctypes.data, ctypes.byref(ps), ctypes.byref(maxp))
        return {k: v[:n].copy() for k, v in ph.items()}, maxp.value
