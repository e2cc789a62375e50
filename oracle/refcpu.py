"""ctypes binding of oracle/_ref/libref_oracle.so - the UNMODIFIED reference sources
(main.cpp, repwvl_thermal.cpp, cplkavg.cpp, lbl.arts/ascii.cpp) compiled from
/root/reference by oracle/Makefile and driven by oracle/ref_harness.cpp.

TEST INFRASTRUCTURE ONLY.  `available()` is False on a checkout where the library
was never built (no /root/reference); callers then fall back to oracle.port.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref_oracle.so")
_lib = None

NLAY, NLEV = 20, 21
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def available() -> bool:
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        L.ref_lowerpos.restype = C.c_long
        L.ref_lowerpos.argtypes = [_dp, C.c_int, C.c_double]
        L.ref_cplkavg.restype = C.c_double
        L.ref_cplkavg.argtypes = [C.c_double] * 3
        L.ref_read_tau.restype = C.c_int
        L.ref_advance.restype = C.c_int
        L.ref_ascii_file2xy2D.restype = C.c_int
        _lib = L
    return _lib


def consts():
    out = np.zeros(8)
    lib().ref_consts(_p(out))
    return dict(zip(["tau_s", "mu_s", "g_asym", "albedo", "daytime", "E_0", "doublings", "cloud_layer"], out))


def solar():
    out = np.zeros(7)
    lib().ref_solar(_p(out))
    return dict(zip(["r_dir", "s_dir", "t_dir", "r", "t", "r_total", "solar_irr"], out))


def lowerpos(a, x):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return int(lib().ref_lowerpos(a, len(a), float(x)))


def init_columns(plevel, Tlevel, vmr_ppm_level, co2_factor=1.0):
    """Tlevel [ncol,21]; vmr_ppm_level [ncol,5,21] (H2O,O3,CO2,CH4,N2O) -> dict of layer state."""
    Tlevel = np.ascontiguousarray(Tlevel, dtype=np.float64).reshape(-1, NLEV)
    ncol = Tlevel.shape[0]
    vm = np.ascontiguousarray(vmr_ppm_level, dtype=np.float64).reshape(ncol, 5, NLEV)
    pl = np.ascontiguousarray(plevel, dtype=np.float64)
    out = dict(Tlayer=np.zeros((ncol, NLAY)), vmr9=np.zeros((ncol, 9, NLAY)), rel_hum=np.zeros((ncol, NLAY)),
               player=np.zeros(NLAY), conv=np.zeros(NLAY))
    lib().ref_init_columns(C.c_int(ncol), _p(pl), _p(Tlevel), _p(vm), C.c_double(co2_factor), _p(out["Tlayer"]),
                           _p(out["vmr9"]), _p(out["rel_hum"]), _p(out["player"]), _p(out["conv"]))
    return out


def read_tau(table, plevel, Tlayer, vmr9, cloud_on=True, nwvl_max=4096):
    pl = np.ascontiguousarray(plevel, dtype=np.float64)
    T = np.ascontiguousarray(Tlayer, dtype=np.float64)
    v = np.ascontiguousarray(vmr9, dtype=np.float64)
    tau = np.zeros((nwvl_max, NLAY)); wvl = np.zeros(nwvl_max); wgt = np.zeros(nwvl_max)
    n = lib().ref_read_tau(table.encode(), _p(pl), _p(T), _p(v), C.c_int(int(cloud_on)), _p(tau), _p(wvl), _p(wgt))
    return tau[:n].copy(), wvl[:n].copy(), wgt[:n].copy()


def radiative_transfer(tau, wvl, weight, Tlayer, T_surface, solar_irr):
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    n = tau.shape[0]
    Ed = np.zeros(NLEV); Eu = np.zeros(NLEV); dE = np.zeros(NLAY)
    lib().ref_radiative_transfer(C.c_int(n), _p(tau), _p(np.ascontiguousarray(wvl, dtype=np.float64)),
                                 _p(np.ascontiguousarray(weight, dtype=np.float64)),
                                 _p(np.ascontiguousarray(Tlayer, dtype=np.float64)), C.c_double(T_surface),
                                 C.c_double(solar_irr), _p(Ed), _p(Eu), _p(dE))
    return Ed, Eu, dE


def advance(table, plevel, rel_hum, solar_irr, Tlayer, Tsurf, vmr9, nsteps, first_step=0, cloud_on=True,
            time_h=None, want_trace=False):
    """Run the reference time loop; returns a dict with the new state and last-step fluxes."""
    Tl = np.array(Tlayer, dtype=np.float64, order="C").reshape(-1, NLAY)
    ncol = Tl.shape[0]
    Ts = np.array(np.broadcast_to(Tsurf, (ncol,)), dtype=np.float64, order="C")
    v9 = np.array(vmr9, dtype=np.float64, order="C").reshape(ncol, 9, NLAY)
    rh = np.ascontiguousarray(rel_hum, dtype=np.float64).reshape(ncol, NLAY)
    pl = np.ascontiguousarray(plevel, dtype=np.float64)
    th = np.zeros(ncol, dtype=np.float32) if time_h is None else np.array(time_h, dtype=np.float32, order="C")
    Ed = np.zeros((ncol, NLEV)); Eu = np.zeros((ncol, NLEV)); dE = np.zeros((ncol, NLAY)); dt = np.zeros(ncol)
    tr = np.zeros((ncol, nsteps, 24)) if want_trace else None
    nw = lib().ref_advance(table.encode(), C.c_int(ncol), C.c_int(first_step), C.c_int(nsteps), _p(pl), _p(rh),
                           C.c_double(solar_irr), C.c_int(int(cloud_on)), _p(Tl), _p(Ts), _p(v9), _p(th), _p(Ed),
                           _p(Eu), _p(dE), _p(dt), _p(tr))
    return dict(Tlayer=Tl, Tsurf=Ts, vmr9=v9, time_h=th, E_down=Ed, E_up=Eu, dE=dE, dt=dt, trace=tr, nwvl=nw)


def cplkavg(lo, hi, t):
    return float(lib().ref_cplkavg(float(lo), float(hi), float(t)))


def cplkavg_many(lo, hi, t):
    lo = np.ascontiguousarray(lo, dtype=np.float64); hi = np.ascontiguousarray(hi, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.float64)
    out = np.zeros_like(lo)
    lib().ref_cplkavg_many(C.c_int(lo.size), _p(lo), _p(hi), _p(t), _p(out))
    return out


def ascii_file2xy2D(path):
    nx = C.c_int(0); ny = C.c_int(0)
    st = lib().ref_ascii_file2xy2D(path.encode(), C.byref(nx), C.byref(ny), None, None)
    if st != 0:
        return st, None, None
    x = np.zeros(nx.value); y = np.zeros((nx.value, ny.value))
    st = lib().ref_ascii_file2xy2D(path.encode(), C.byref(nx), C.byref(ny), _p(x), _p(y))
    return st, x, y
