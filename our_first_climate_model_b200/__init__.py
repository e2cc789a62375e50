"""B200-native radiative-convective column solver (thermal hot path of
pabloconrat/our_first_climate_model).

The product is the C-ABI shared library `lib/librcm_b200.so` (CUDA kernels for sm_100a +
host C++, sources under `csrc/`, interface in `include/rcm_b200.h`).  This package is the thin
ctypes binding used by the tests, the bench and Python drivers; it contains no numerics and
no CPU fallback - every compute call fails loudly when the library or a GPU is missing.
"""
from .capi import (RcmError, Solver, Table, StepScalars, default_params, default_solar_params, solar_setup,  # noqa: F401
                   lowerpos, read_atm, init_columns, make_ensemble, make_lbl_tables, write_lbl_asc, ascii_file2xy2D, cplkavg_host, device_count,
                   write_profiles, SolarParams, library_path, load_library, build_library, DECLARED_SYMBOLS)

__all__ = ["RcmError", "Solver", "Table", "StepScalars", "default_params", "default_solar_params", "solar_setup",
           "lowerpos", "read_atm", "init_columns", "make_ensemble", "make_lbl_tables", "write_lbl_asc", "ascii_file2xy2D", "cplkavg_host", "device_count",
           "write_profiles", "SolarParams", "library_path", "load_library", "build_library", "DECLARED_SYMBOLS"]
