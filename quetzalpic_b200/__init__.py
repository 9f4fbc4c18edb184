"""quetzalpic_b200 -- B200-native (sm_100a) per-xi-slice quasistatic wakefield solver.

Drop-in for the hot path of QuetzalPIC, `System_3d::solve_wakefield` (reference src/System_3d.cpp:238-627):
CUDA kernels + C ABI in csrc/ (built to libqpic_b200.so), `solver.py` = ctypes binding and host-side mirror of the
reference's solver state.  CUDA only -- there is no CPU fallback.
"""
from .solver import (  # noqa: F401
    ABI_SYMBOLS, FIELDS_2D, FIELDS_3D, LIBRARY_PATH, PARTICLE_FIELDS, SLICE_FIELDS_3D, QpicError, QpicLibrary,
    WakefieldSolver, gaussian_laser_widths, grid_dimensions, load_library, powerlaw2d, wake_particle_lattice,
    BUNCH_FIELDS, SPECIES, SPECIES_IDS, bunch_particle_lattice, bunch_profile_from_table, concatenate_bunches, cylindrical3d, gaussian3d,
    parabolic3d, plasma_units,
)
