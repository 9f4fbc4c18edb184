"""ctypes binding of the qpic_b200 C ABI (include/qpic_b200.h) and a host-side mirror of the reference's
solver interface.

The reference is a C++ program without a plugin boundary; its hot path is `System_3d::solve_wakefield`
(reference src/System_3d.cpp:238-627) operating on `System_3d`'s members.  `WakefieldSolver` below exposes
the same state under the reference's names (psi, rho, jx, jy, jz, by, bz, ex, ey, ez, aSqr, susceptibility;
src/System_3d.cpp:49-62) and the same call order (init_wake_particles -> solve_wakefield), so the parity
tests read like a transcript of `System_3d::run()` (src/System_3d.cpp:199-236).

There is NO CPU fallback here: the shared library is CUDA-only and loading / creating a context fails loudly
without a B200-class device.
"""
import ctypes
import os

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# QPIC_B200_LIBRARY: development override (A/B builds of the same sources)
LIBRARY_PATH = os.environ.get("QPIC_B200_LIBRARY") or os.path.join(_PKG_DIR, "libqpic_b200.so")

FIELDS_2D = ["psi", "psi_prev", "psi_middle", "rho", "jx", "jy", "jz", "jy_next", "jz_next", "rho_ion", "by_cur", "bz_cur"]
FIELDS_3D = ["aSqr", "susceptibility", "rho_bunch", "jx_bunch", "ex", "ey", "ez", "by", "bz"]
SLICE_FIELDS_3D = ["psi", "rho", "jx", "jy", "jz"]  # filled slice by slice, like Output_writer::write_slice
PARTICLE_FIELDS = ["y", "z", "py", "pz", "n", "gamma", "px"]
BUNCH_FIELDS = ["x", "y", "z", "px", "py", "pz", "gamma", "n", "ex", "ey", "ez", "by", "bz", "chi"]  # bunch_particle_3d
# built-in species, ids 0-3 (reference src/Config_reader.cpp:19-22): name -> (charge, mass, is_photon)
SPECIES = {"electron": (-1.0, 1.0, False), "positron": (1.0, 1.0, False), "proton": (1.0, 1836.15267343, False), "photon": (0.0, 0.0, True)}
SPECIES_IDS = {name: i for i, name in enumerate(SPECIES)}

# every symbol include/qpic_b200.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "qpic_default_params", "qpic_create", "qpic_destroy", "qpic_last_error", "qpic_abi_version", "qpic_source_hash",
    "qpic_set_laser_gaussian", "qpic_upload_a_sqr", "qpic_upload_bunch_sources", "qpic_init_wake_particles",
    "qpic_solve_wakefield", "qpic_begin", "qpic_step", "qpic_finish", "qpic_synchronize", "qpic_current_slice",
    "qpic_stopped_at", "qpic_download_field3d", "qpic_download_slice3d", "qpic_download_field2d",
    "qpic_upload_field2d", "qpic_particle_count", "qpic_download_particles", "qpic_upload_particles",
    "qpic_set_current_slice", "qpic_solve_poisson_host", "qpic_solve_helmholtz_pair_host", "qpic_last_step_stats",
    "qpic_profile_slices", "qpic_upload_a_sqr_slices", "qpic_download_slices3d", "qpic_alloc_pinned", "qpic_free_pinned",
    "qpic_fast_path_aborts", "qpic_wait_transfers", "qpic_timeline_reset", "qpic_timeline_read", "qpic_timeline_kernel_name",
    "qpic_dist_region_bytes", "qpic_dist_region", "qpic_dist_ipc_handle", "qpic_dist_attach_ipc", "qpic_dist_attach_pointers", "qpic_dist_row_range", "qpic_download_xy_cut",
    "qpic_bunch_set_species", "qpic_bunch_init", "qpic_bunch_count", "qpic_bunch_deposit_densities", "qpic_bunch_gather_fields",
    "qpic_bunch_advance", "qpic_bunch_download", "qpic_bunch_upload", "qpic_bunch_advance_qed", "qpic_photon_count", "qpic_photons_download",
]


class QpicError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"qpic_b200 error {code}: {message}")
        self.code = code


class QpicParams(ctypes.Structure):
    _fields_ = [
        ("nx", ctypes.c_int), ("ny", ctypes.c_int), ("nz", ctypes.c_int),
        ("lx", ctypes.c_double), ("ly", ctypes.c_double), ("lz", ctypes.c_double),
        ("magnetic_field_iterations", ctypes.c_int),
        ("psi_threshold", ctypes.c_double),
        ("magnetic_field_D", ctypes.c_double),
        ("threshold_B", ctypes.c_double),
        ("device", ctypes.c_int),
        ("keep_slices_3d", ctypes.c_int),
        ("use_graph", ctypes.c_int),
        ("stop_check_interval", ctypes.c_int),
        ("force_general_deposit", ctypes.c_int),
        ("dist_rank", ctypes.c_int), ("dist_world", ctypes.c_int),
        ("particle_capacity", ctypes.c_size_t),
        ("keep_xy_cuts", ctypes.c_int), ("xy_z0", ctypes.c_double),
    ]


class QpicLibrary:
    """A loaded qpic_b200 shared library with typed entry points."""

    def __init__(self, path=LIBRARY_PATH):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path = path
        L = ctypes.CDLL(path)
        vp, dbl, ci = ctypes.c_void_p, ctypes.c_double, ctypes.c_int
        L.qpic_default_params.argtypes = [ctypes.POINTER(QpicParams)]
        L.qpic_default_params.restype = None
        L.qpic_create.argtypes = [ctypes.POINTER(QpicParams), ctypes.POINTER(vp)]
        L.qpic_destroy.argtypes = [vp]
        L.qpic_destroy.restype = None
        L.qpic_last_error.argtypes = [vp]
        L.qpic_last_error.restype = ctypes.c_char_p
        L.qpic_set_laser_gaussian.argtypes = [vp] + [dbl] * 7
        L.qpic_upload_a_sqr.argtypes = [vp, vp]
        L.qpic_upload_bunch_sources.argtypes = [vp, vp, vp]
        L.qpic_init_wake_particles.argtypes = [vp, vp, vp, vp, ctypes.c_size_t]
        L.qpic_solve_wakefield.argtypes = [vp]
        L.qpic_begin.argtypes = [vp]
        L.qpic_step.argtypes = [vp, ci, ctypes.POINTER(ci)]
        L.qpic_finish.argtypes = [vp]
        L.qpic_synchronize.argtypes = [vp]
        L.qpic_current_slice.argtypes = [vp]
        L.qpic_stopped_at.argtypes = [vp]
        L.qpic_download_field3d.argtypes = [vp, ci, vp]
        L.qpic_download_slice3d.argtypes = [vp, ci, ci, vp]
        L.qpic_download_field2d.argtypes = [vp, ci, vp]
        L.qpic_upload_field2d.argtypes = [vp, ci, vp]
        L.qpic_particle_count.argtypes = [vp]
        L.qpic_particle_count.restype = ctypes.c_size_t
        L.qpic_download_particles.argtypes = [vp, ci, vp]
        L.qpic_upload_particles.argtypes = [vp, ci, vp]
        L.qpic_set_current_slice.argtypes = [vp, ci]
        L.qpic_solve_poisson_host.argtypes = [vp, vp, dbl]
        L.qpic_solve_helmholtz_pair_host.argtypes = [vp, vp, vp, dbl]
        L.qpic_last_step_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_long)]
        L.qpic_profile_slices.argtypes = [vp, ci, ctypes.c_char_p, ctypes.c_size_t]
        L.qpic_upload_a_sqr_slices.argtypes = [vp, ci, ci, vp]
        L.qpic_download_slices3d.argtypes = [vp, ci, ci, ci, vp]
        L.qpic_alloc_pinned.argtypes = [ctypes.c_size_t]
        L.qpic_alloc_pinned.restype = vp
        L.qpic_free_pinned.argtypes = [vp]
        L.qpic_free_pinned.restype = None
        L.qpic_wait_transfers.argtypes = [vp]
        L.qpic_fast_path_aborts.argtypes = [vp]
        L.qpic_fast_path_aborts.restype = ctypes.c_long
        L.qpic_download_xy_cut.argtypes = [vp, ci, ci, ci, vp]
        L.qpic_dist_region_bytes.argtypes = [vp]
        L.qpic_dist_region_bytes.restype = ctypes.c_size_t
        L.qpic_dist_region.argtypes = [vp]
        L.qpic_dist_region.restype = vp
        L.qpic_dist_ipc_handle.argtypes = [vp, vp]
        L.qpic_dist_attach_ipc.argtypes = [vp, vp]
        L.qpic_dist_attach_pointers.argtypes = [vp, ctypes.POINTER(vp)]
        L.qpic_dist_row_range.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ci)]
        L.qpic_bunch_set_species.argtypes = [vp, ci, vp, vp, vp]
        L.qpic_bunch_init.argtypes = [vp, ctypes.c_size_t] + [vp] * 9
        L.qpic_bunch_count.argtypes = [vp]
        L.qpic_bunch_count.restype = ctypes.c_size_t
        L.qpic_bunch_deposit_densities.argtypes = [vp]
        L.qpic_bunch_gather_fields.argtypes = [vp, dbl]
        L.qpic_bunch_advance.argtypes = [vp, dbl]
        L.qpic_bunch_download.argtypes = [vp, ci, vp]
        L.qpic_bunch_advance_qed.argtypes = [vp, dbl, dbl, vp, vp, ctypes.c_ulonglong, ctypes.POINTER(dbl), ctypes.POINTER(ctypes.c_size_t)]
        L.qpic_photon_count.argtypes = [vp]
        L.qpic_photon_count.restype = ctypes.c_size_t
        L.qpic_photons_download.argtypes = [vp, ci, vp]
        L.qpic_bunch_upload.argtypes = [vp, ci, vp]
        self.L = L

    def pinned_array(self, shape):
        """float64 numpy array backed by page-locked host memory (freed with free_pinned)."""
        count = int(np.prod(shape))
        ptr = self.L.qpic_alloc_pinned(count * 8)
        if not ptr:
            raise MemoryError("cudaMallocHost failed")
        buf = (ctypes.c_double * count).from_address(ptr)
        arr = np.frombuffer(buf, dtype=np.float64).reshape(shape)
        return arr, ptr

    def free_pinned(self, ptr):
        self.L.qpic_free_pinned(ptr)


_product_library = None

CSRC_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
ABI_HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "qpic_b200.h")


def source_hash():
    """sha256 (first 16 hex digits) over the CUDA sources and the ABI header: what __graft_entry__.build() embeds in the library"""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR) if f.endswith((".cu", ".cuh", ".h"))) + [ABI_HEADER]
    for path in files:
        h.update(os.path.basename(path).encode())
        h.update(open(path, "rb").read())
    return h.hexdigest()[:16]


def load_library():
    """The product library, quetzalpic_b200/libqpic_b200.so (CUDA, sm_100a).  Refuses a binary that was not built from the sources
    next to it (the .so is git-ignored and travels with the tree: a stale one must not be used silently)."""
    global _product_library
    if _product_library is None:
        lib = QpicLibrary(LIBRARY_PATH)
        lib.L.qpic_source_hash.restype = ctypes.c_char_p
        built_from = lib.L.qpic_source_hash().decode()
        if os.path.isdir(CSRC_DIR) and built_from != source_hash():
            raise RuntimeError(f"{LIBRARY_PATH} was built from other sources (embedded hash {built_from}, sources {source_hash()}): "
                               "rebuild it with python -c 'import __graft_entry__ as g; g.build()'")
        _product_library = lib
    return _product_library


def grid_dimensions(l, d):
    """n = static_cast<int>(l / d) per axis (reference src/System_3d.cpp:87-89; note 4.8/0.2 -> 23)."""
    return tuple(int(float(a) / float(b)) for a, b in zip(l, d))


def wake_particle_lattice(ny, nz, ly, lz, ppcy, ppcz, profile=None, rho0=-1.0):
    """Positions and weights of System_3d::init_wake_particles (reference src/System_3d.cpp:724-748).

    `profile(y, z)` (vectorised) gives the plasma density; None = constant2d(rho0) (src/profiles.cpp:47-51).
    Particles where the profile is exactly 0 are skipped (:735).  Order: y-major, then z (:730-731).
    """
    dy, dz = ly / ny, lz / nz
    yy = (np.arange(ppcy * ny) + 0.5) * dy / ppcy
    zz = (np.arange(ppcz * nz) + 0.5) * dz / ppcz
    y, z = np.meshgrid(yy, zz, indexing="ij")
    y, z = y.ravel(), z.ravel()
    value = np.full(y.shape, float(rho0)) if profile is None else np.asarray(profile(y, z), dtype=np.float64)
    keep = value != 0
    return y[keep].copy(), z[keep].copy(), (value[keep] / ppcy / ppcz).copy()


def powerlaw2d(power, radius, y0, z0, min_factor=0.0, max_factor=1e10, amplitude=-1.0):
    """Plasma channel profile, reference src/profiles.cpp:53-64 (Config_reader.cpp:243-252 passes amplitude -1)."""
    def profile(y, z):
        r = np.sqrt((y - y0) * (y - y0) + (z - z0) * (z - z0))
        value = np.power(r / radius, power)
        value = np.minimum(value, max_factor)
        value = np.maximum(value, min_factor)
        return amplitude * value
    return profile


def plasma_units(base_frequency_SI=1.0):
    """Plasma_units (reference src/Plasma_units.h:16-33, constants src/constants.h:3-10): the two derived quantities the beam
    path needs -- field_schwinger (chi) and number_density_norm (the macro-particle weight written to Bunch_NNN.h5)."""
    pi = 3.141592653589793238
    electron_mass, planck, c_light, e_charge = 9.1093837015e-28, 6.62607015e-27, 29979245800.0, 4.803204672997e-10
    hbar = planck / (2 * pi)
    frequency = float(base_frequency_SI)
    wavenumber = frequency / c_light
    density = electron_mass * frequency * frequency / (4 * pi * e_charge * e_charge)
    return dict(frequency=frequency, plasma_wavenumber=wavenumber, wavelength=2 * pi / wavenumber, density=density,
                number_density_norm=density / wavenumber / wavenumber / wavenumber,
                field_schwinger=electron_mass * c_light * c_light / hbar / frequency)


def gaussian3d(amplitude, width, r0):
    """reference src/profiles.cpp:4-11 (vectorised)"""
    def profile(x, y, z):
        return amplitude * np.exp(-(x - r0[0]) * (x - r0[0]) / width[0] / width[0] - (y - r0[1]) * (y - r0[1]) / width[1] / width[1]
                                  - (z - r0[2]) * (z - r0[2]) / width[2] / width[2])
    return profile


def parabolic3d(amplitude, xsize, rsize, r0):
    """reference src/profiles.cpp:19-31"""
    def profile(x, y, z):
        r_squared = (y - r0[1]) * (y - r0[1]) + (z - r0[2]) * (z - r0[2])
        x_rel_squared = (x - r0[0]) * (x - r0[0])
        inside = (r_squared <= rsize * rsize) & (x_rel_squared < xsize * xsize)
        return np.where(inside, amplitude * (1 - x_rel_squared / xsize / xsize) * (1 - r_squared / rsize / rsize), 0.0)
    return profile


def cylindrical3d(amplitude, xsize, rsize, r0):
    """reference src/profiles.cpp:33-45"""
    def profile(x, y, z):
        r_squared = (y - r0[1]) * (y - r0[1]) + (z - r0[2]) * (z - r0[2])
        x_rel_squared = (x - r0[0]) * (x - r0[0])
        return np.where((r_squared <= rsize * rsize) & (x_rel_squared < xsize * xsize), amplitude, 0.0)
    return profile


def bunch_profile_from_table(table, l):
    """The density profile Config_reader::parse_bunch_table builds from a [bunch] table (reference src/Config_reader.cpp:147-182)."""
    shape = table["shape"]
    y0, z0 = float(table.get("y0", 0.5 * l[1])), float(table.get("z0", 0.5 * l[2]))
    if shape == "gaussian":
        s = np.sqrt(8.0)
        width = (table["xsigma"] / s, table["ysigma"] / s, table["zsigma"] / s)
        return gaussian3d(float(table["rho0"]), width, (float(table.get("x0", table["xsigma"])), y0, z0))
    if shape == "parabolic":
        return parabolic3d(float(table["rho0"]), float(table["xsize"]), float(table["rsize"]), (float(table["x0"]), y0, z0))
    if shape == "cylindrical":
        return cylindrical3d(float(table["rho0"]), float(table["xsize"]), float(table["rsize"]), (float(table["x0"]), y0, z0))
    raise ValueError(f'Bunch shape "{shape}" is not supported, use "gaussian", "parabolic", or "cylindrical".')


def bunch_particle_lattice(n, l, ppc, rho, gamma, species="electron"):
    """The particles System_3d::init_bunch_particles creates for one bunch (reference src/System_3d.cpp:767-803): a regular
    lattice of ppc sub-cell points, kept where rho(x, y, z) > 0, x-major order; px = sqrt(gamma^2 - 1) (photons: gamma).
    Returns a dict of the BUNCH_FIELDS that are set plus "species_id"."""
    nx, ny, nz = n
    dx, dy, dz = l[0] / nx, l[1] / ny, l[2] / nz
    xx = (np.arange(ppc[0] * nx) + 0.5) * dx / ppc[0]
    yy = (np.arange(ppc[1] * ny) + 0.5) * dy / ppc[1]
    zz = (np.arange(ppc[2] * nz) + 0.5) * dz / ppc[2]
    x, y, z = (a.ravel() for a in np.meshgrid(xx, yy, zz, indexing="ij"))
    value = np.asarray(rho(x, y, z), dtype=np.float64)
    keep = value > 0.0
    count = int(keep.sum())
    sid = SPECIES_IDS[species]
    px0 = gamma if SPECIES[species][2] else np.sqrt(gamma * gamma - 1)
    return dict(x=x[keep].copy(), y=y[keep].copy(), z=z[keep].copy(), px=np.full(count, px0), py=np.zeros(count), pz=np.zeros(count),
                gamma=np.full(count, float(gamma)), n=(value[keep] / ppc[0] / ppc[1] / ppc[2]).copy(),
                species_id=np.full(count, sid, dtype=np.int32))


def concatenate_bunches(bunches):
    """bunch_particles of System_3d: the bunches one after another (reference src/System_3d.cpp:157-162)"""
    keys = ["x", "y", "z", "px", "py", "pz", "gamma", "n", "species_id"]
    return {k: np.concatenate([b[k] for b in bunches]) if bunches else np.zeros(0, dtype=np.int32 if k == "species_id" else np.float64)
            for k in keys}


def gaussian_laser_widths(xsigma, ysigma, zsigma):
    """Full width at 1/e^2 -> Gaussian parameter, reference src/Config_reader.cpp:124-126."""
    s = np.sqrt(8.0)
    return xsigma / s, ysigma / s, zsigma / s


class WakefieldSolver:
    """Device-resident quasistatic wakefield solver: the state System_3d::solve_wakefield works on."""

    def __init__(self, nx, ny, nz, lx, ly, lz, magnetic_field_iterations, psi_threshold=1e-4, magnetic_field_D=10.0,
                 threshold_B=100.0, device=0, keep_slices_3d=False, use_graph=True, stop_check_interval=64, force_general_deposit=False,
                 library=None, dist_rank=0, dist_world=1, particle_capacity=0, keep_xy_cuts=False, xy_z0=None):
        self.lib = library if library is not None else load_library()
        L = self.lib.L
        p = QpicParams()
        L.qpic_default_params(ctypes.byref(p))
        p.nx, p.ny, p.nz = int(nx), int(ny), int(nz)
        p.lx, p.ly, p.lz = float(lx), float(ly), float(lz)
        p.magnetic_field_iterations = int(magnetic_field_iterations)
        p.psi_threshold, p.magnetic_field_D, p.threshold_B = float(psi_threshold), float(magnetic_field_D), float(threshold_B)
        p.device = int(device)
        p.keep_slices_3d = int(bool(keep_slices_3d))
        p.use_graph = int(bool(use_graph))
        p.stop_check_interval = int(stop_check_interval)
        p.force_general_deposit = int(bool(force_general_deposit))
        p.dist_rank, p.dist_world, p.particle_capacity = int(dist_rank), int(dist_world), int(particle_capacity)
        self.dist_rank, self.dist_world = int(dist_rank), int(dist_world)
        p.keep_xy_cuts = int(bool(keep_xy_cuts))
        p.xy_z0 = -1.0 if xy_z0 is None else float(xy_z0)
        self.params = p
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.keep_slices_3d = bool(keep_slices_3d)
        handle = ctypes.c_void_p()
        rc = L.qpic_create(ctypes.byref(p), ctypes.byref(handle))
        if rc != 0:
            raise QpicError(rc, L.qpic_last_error(None).decode())
        self.ctx = handle

    # -- plumbing
    def _check(self, rc):
        if rc != 0:
            raise QpicError(rc, self.lib.L.qpic_last_error(self.ctx).decode())

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.L.qpic_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- inputs
    def set_laser_gaussian(self, amplitude, wx, wy, wz, x0, y0, z0):
        self._check(self.lib.L.qpic_set_laser_gaussian(self.ctx, amplitude, wx, wy, wz, x0, y0, z0))

    def upload_a_sqr(self, a_sqr):
        a = np.ascontiguousarray(a_sqr, dtype=np.float64)
        if a.shape != (self.nx, self.ny, self.nz):
            raise ValueError("a_sqr must have shape (nx, ny, nz)")
        self._check(self.lib.L.qpic_upload_a_sqr(self.ctx, a.ctypes.data))

    def upload_bunch_sources(self, rho_bunch=None, jx_bunch=None):
        arrs = []
        for a in (rho_bunch, jx_bunch):
            if a is None:
                arrs.append(None)
            else:
                a = np.ascontiguousarray(a, dtype=np.float64)
                if a.shape != (self.nx, self.ny, self.nz):
                    raise ValueError("bunch sources must have shape (nx, ny, nz)")
                arrs.append(a)
        self._check(self.lib.L.qpic_upload_bunch_sources(
            self.ctx, None if arrs[0] is None else arrs[0].ctypes.data, None if arrs[1] is None else arrs[1].ctypes.data))

    def init_wake_particles(self, y, z, n):
        y = np.ascontiguousarray(y, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.float64)
        n = np.ascontiguousarray(n, dtype=np.float64)
        if not (y.shape == z.shape == n.shape and y.ndim == 1):
            raise ValueError("y, z, n must be 1-D arrays of equal length")
        self._check(self.lib.L.qpic_init_wake_particles(self.ctx, y.ctypes.data, z.ctypes.data, n.ctypes.data, y.size))

    # -- sweep
    def solve_wakefield(self):
        self._check(self.lib.L.qpic_solve_wakefield(self.ctx))
        return self.stopped_at

    def begin(self):
        self._check(self.lib.L.qpic_begin(self.ctx))

    def step(self, max_slices=1, wait=True):
        """Runs up to max_slices xi-slices.  wait=True returns whether the sweep is over (synchronises);
        wait=False only queues the work."""
        if wait:
            done = ctypes.c_int(0)
            self._check(self.lib.L.qpic_step(self.ctx, int(max_slices), ctypes.byref(done)))
            return bool(done.value)
        self._check(self.lib.L.qpic_step(self.ctx, int(max_slices), None))
        return None

    def finish(self):
        self._check(self.lib.L.qpic_finish(self.ctx))

    def synchronize(self):
        self._check(self.lib.L.qpic_synchronize(self.ctx))

    @property
    def current_slice(self):
        return self.lib.L.qpic_current_slice(self.ctx)

    @current_slice.setter
    def current_slice(self, value):
        self._check(self.lib.L.qpic_set_current_slice(self.ctx, int(value)))

    @property
    def stopped_at(self):
        return self.lib.L.qpic_stopped_at(self.ctx)

    @property
    def fast_path_aborts(self):
        return self.lib.L.qpic_fast_path_aborts(self.ctx)

    def last_step_stats(self):
        ms = ctypes.c_float(0)
        launches = ctypes.c_long(0)
        self._check(self.lib.L.qpic_last_step_stats(self.ctx, ctypes.byref(ms), ctypes.byref(launches)))
        return ms.value, launches.value

    def profile_slices(self, nslices=1):
        """Eager slices with an event after every kernel: {kernel: (launches, total_ms)} in first-launch order."""
        buf = ctypes.create_string_buffer(16384)
        self._check(self.lib.L.qpic_profile_slices(self.ctx, int(nslices), buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, count, ms = line.split()
            out[name] = (int(count), float(ms))
        return out

    def upload_a_sqr_slices(self, first_slice, a_sqr_block):
        """asynchronous; a_sqr_block: (count, ny, nz) contiguous float64 (ideally pinned)"""
        assert a_sqr_block.flags["C_CONTIGUOUS"] and a_sqr_block.dtype == np.float64
        self._check(self.lib.L.qpic_upload_a_sqr_slices(self.ctx, int(first_slice), int(a_sqr_block.shape[0]), a_sqr_block.ctypes.data))

    def download_slices3d(self, name, first_slice, out_block, per_slice=False):
        """asynchronous; out_block: (count, ny, nz) contiguous float64 (ideally pinned); synchronize() before reading"""
        assert out_block.flags["C_CONTIGUOUS"] and out_block.dtype == np.float64
        self._check(self.lib.L.qpic_download_slices3d(self.ctx, self._field3d_index(name, per_slice), int(first_slice),
                                                      int(out_block.shape[0]), out_block.ctypes.data))

    def download_xy_cut(self, name, first_slice, out_block, per_slice=False):
        """asynchronous; out_block: (count, ny) contiguous float64: the Fields_xy cut (z = xy_z0) of `count` slices"""
        assert out_block.flags["C_CONTIGUOUS"] and out_block.dtype == np.float64 and out_block.shape[1] == self.ny
        self._check(self.lib.L.qpic_download_xy_cut(self.ctx, self._field3d_index(name, per_slice), int(first_slice), int(out_block.shape[0]),
                                                    out_block.ctypes.data))

    def xy_cut(self, name, per_slice=False):
        """(nx, ny) Fields_xy dataset of a field (reference src/array_utils.cpp:7-24)"""
        out = np.empty((self.nx, self.ny))
        self.download_xy_cut(name, 0, out, per_slice=per_slice)
        self.wait_transfers()
        return out

    def wait_transfers(self):
        self._check(self.lib.L.qpic_wait_transfers(self.ctx))

    # -- state
    def field2d(self, name):
        out = np.empty((self.ny, self.nz))
        self._check(self.lib.L.qpic_download_field2d(self.ctx, FIELDS_2D.index(name), out.ctypes.data))
        return out

    def set_field2d(self, name, value):
        a = np.ascontiguousarray(value, dtype=np.float64)
        if a.shape != (self.ny, self.nz):
            raise ValueError("2-D fields have shape (ny, nz)")
        self._check(self.lib.L.qpic_upload_field2d(self.ctx, FIELDS_2D.index(name), a.ctypes.data))

    def _field3d_index(self, name, per_slice):
        if per_slice:
            return len(FIELDS_3D) + SLICE_FIELDS_3D.index(name)
        return FIELDS_3D.index(name)

    def field3d(self, name, per_slice=False):
        """(nx, ny, nz) array; per_slice=True selects the slice-by-slice accumulation of psi/rho/jx/jy/jz."""
        out = np.empty((self.nx, self.ny, self.nz))
        self._check(self.lib.L.qpic_download_field3d(self.ctx, self._field3d_index(name, per_slice), out.ctypes.data))
        return out

    def slice3d(self, name, i, per_slice=False):
        out = np.empty((self.ny, self.nz))
        self._check(self.lib.L.qpic_download_slice3d(self.ctx, self._field3d_index(name, per_slice), int(i), out.ctypes.data))
        return out

    @property
    def particle_count(self):
        return self.lib.L.qpic_particle_count(self.ctx)

    def particles(self, name):
        out = np.empty(self.particle_count)
        self._check(self.lib.L.qpic_download_particles(self.ctx, PARTICLE_FIELDS.index(name), out.ctypes.data))
        return out

    def set_particles(self, name, value):
        a = np.ascontiguousarray(value, dtype=np.float64)
        if a.shape != (self.particle_count,):
            raise ValueError("particle arrays have shape (particle_count,)")
        self._check(self.lib.L.qpic_upload_particles(self.ctx, PARTICLE_FIELDS.index(name), a.ctypes.data))

    # -- beam particles: the part of System_3d::run around solve_wakefield (reference src/System_3d.cpp:199-236)
    def bunch_init(self, particles, species=None):
        """particles: dict with x, y, z, px, py, pz, gamma, n, species_id (see bunch_particle_lattice / concatenate_bunches);
        species: list of (charge, mass, is_photon), default the four built-ins"""
        table = list(SPECIES.values()) if species is None else list(species)
        charge = np.array([s[0] for s in table], dtype=np.float64)
        cmr = np.array([0.0 if s[1] == 0.0 else s[0] / s[1] for s in table], dtype=np.float64)  # src/Species_register.h:12-13
        photon = np.array([1 if s[2] else 0 for s in table], dtype=np.int32)
        self._check(self.lib.L.qpic_bunch_set_species(self.ctx, len(table), charge.ctypes.data, cmr.ctypes.data, photon.ctypes.data))
        arrs = [np.ascontiguousarray(particles[k], dtype=np.float64) for k in ["x", "y", "z", "px", "py", "pz", "gamma", "n"]]
        sid = np.ascontiguousarray(particles["species_id"], dtype=np.int32)
        if any(a.shape != (sid.size,) for a in arrs):
            raise ValueError("beam particle arrays must be 1-D and of equal length")
        self._check(self.lib.L.qpic_bunch_init(self.ctx, sid.size, *[a.ctypes.data for a in arrs], sid.ctypes.data))

    @property
    def bunch_count(self):
        return self.lib.L.qpic_bunch_count(self.ctx)

    def deposit_bunch_densities(self):
        self._check(self.lib.L.qpic_bunch_deposit_densities(self.ctx))

    def deposit_wakefield_to_bunch(self, field_schwinger):
        self._check(self.lib.L.qpic_bunch_gather_fields(self.ctx, float(field_schwinger)))

    def advance_bunch_particles(self, dt):
        self._check(self.lib.L.qpic_bunch_advance(self.ctx, float(dt)))

    def advance_bunch_particles_qed(self, dt, frequency, u_emit=None, u_accept=None, seed=0):
        """qed = true (reference src/System_3d.cpp:887-921).  u_emit / u_accept: the two uniform numbers per particle the reference
        takes from rand(), or None = device-side Philox keyed by `seed`.  Returns (photons dict, max radiation probability)."""
        maxp, count = ctypes.c_double(0.0), ctypes.c_size_t(0)
        if u_emit is None:
            p1 = p2 = None
        else:
            u1 = np.ascontiguousarray(u_emit, dtype=np.float64)
            u2 = np.ascontiguousarray(u_accept, dtype=np.float64)
            if not (u1.shape == u2.shape == (self.bunch_count,)):
                raise ValueError("u_emit, u_accept must have shape (bunch_count,)")
            p1, p2 = u1.ctypes.data, u2.ctypes.data
        self._check(self.lib.L.qpic_bunch_advance_qed(self.ctx, float(dt), float(frequency), p1, p2, int(seed), ctypes.byref(maxp), ctypes.byref(count)))
        photons = {}
        for name in BUNCH_FIELDS:
            out = np.empty(count.value)
            self._check(self.lib.L.qpic_photons_download(self.ctx, BUNCH_FIELDS.index(name), out.ctypes.data))
            photons[name] = out
        return photons, maxp.value

    def bunch(self, name):
        out = np.empty(self.bunch_count)
        self._check(self.lib.L.qpic_bunch_download(self.ctx, BUNCH_FIELDS.index(name), out.ctypes.data))
        return out

    def set_bunch(self, name, value):
        a = np.ascontiguousarray(value, dtype=np.float64)
        if a.shape != (self.bunch_count,):
            raise ValueError("beam particle arrays have shape (bunch_count,)")
        self._check(self.lib.L.qpic_bunch_upload(self.ctx, BUNCH_FIELDS.index(name), a.ctypes.data))

    # -- slab-decomposed multi-GPU sweep (see quetzalpic_b200/dist.py)
    @property
    def row_range(self):
        """(first own row, number of own rows): the whole grid on a single GPU"""
        r0, n = ctypes.c_int(0), ctypes.c_int(0)
        self._check(self.lib.L.qpic_dist_row_range(self.ctx, ctypes.byref(r0), ctypes.byref(n)))
        return r0.value, n.value

    def dist_ipc_handle(self):
        buf = ctypes.create_string_buffer(64)
        self._check(self.lib.L.qpic_dist_ipc_handle(self.ctx, buf))
        return buf.raw

    def dist_attach_ipc(self, handles):
        """handles: the 64-byte handles of all ranks, in rank order"""
        blob = b"".join(handles)
        assert len(blob) == 64 * self.dist_world
        self._check(self.lib.L.qpic_dist_attach_ipc(self.ctx, blob))

    def dist_region(self):
        return self.lib.L.qpic_dist_region(self.ctx)

    def dist_attach_pointers(self, regions):
        arr = (ctypes.c_void_p * self.dist_world)(*regions)
        self._check(self.lib.L.qpic_dist_attach_pointers(self.ctx, arr))

    # -- building blocks
    def solve_poisson(self, src, D=0.0):
        a = np.array(src, dtype=np.float64, order="C")
        self._check(self.lib.L.qpic_solve_poisson_host(self.ctx, a.ctypes.data, float(D)))
        return a

    def solve_helmholtz_pair(self, src_a, src_b, D):
        a = np.array(src_a, dtype=np.float64, order="C")
        b = np.array(src_b, dtype=np.float64, order="C")
        self._check(self.lib.L.qpic_solve_helmholtz_pair_host(self.ctx, a.ctypes.data, b.ctypes.data, float(D)))
        return a, b
