"""Checkpoint / restart of an ensemble: a run that is saved, torn down and loaded into a fresh solver
continues bit-identically - for the repwvl path (with per-column forcing) and the line-by-line path."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, table_path

pytestmark = pytest.mark.gpu
KEYS = ("Tlayer", "Tsurf", "h2o", "time_h", "E_down", "E_up", "dE", "dt")


def same(a, b):
    return all(np.array_equal(a[k], b[k]) for k in KEYS)


def test_repwvl_restart_is_bit_identical(rcm, tmp_path):
    atm = rcm.read_atm(os.path.join(GOLDEN, "column21.atm"))
    pl = atm[:, 1].copy()
    n = 53
    Tlev, vlev = rcm.make_ensemble(n, 31, pl, atm[:, 2].copy(), atm[:, 4:9].T.copy())
    st = rcm.init_columns(pl, Tlev, vlev)
    rng = np.random.default_rng(1)
    a = rcm.Solver(0)
    a.set_repwvl_table_from(rcm.Table(table_path(20)))
    a.set_columns(pl, st["Tlayer"], Tlev[:, 20].copy(), st["vmr9"], st["rel_hum"])
    a.set_column_solar(None, rng.uniform(1, 3, n), rng.uniform(0.4, 0.6, n), rng.uniform(0.1, 0.2, n), cloud_from_tau_s=True)
    a.advance(4)
    path = str(tmp_path / "ens.ckpt")
    a.save_checkpoint(path)
    saved = a.get_state()
    sc_a = a.advance(6)
    end_a = a.get_state()
    a.close()
    b = rcm.Solver(0)
    b.set_repwvl_table_from(rcm.Table(table_path(20)))
    b.load_checkpoint(path)
    assert b.ncol == n and same(b.get_state(), saved)
    sc_b = b.advance(6)
    assert same(b.get_state(), end_a) and np.array_equal(sc_a, sc_b)
    # a checkpoint taken before the first step keeps "tau from the unsorted initial profile" (main.cpp:500-504)
    b.set_columns(pl, st["Tlayer"], Tlev[:, 20].copy(), st["vmr9"], st["rel_hum"])
    b.save_checkpoint(path)
    b.advance(2)
    ref = b.get_state()
    b.load_checkpoint(path)
    b.advance(2)
    assert same(b.get_state(), ref)
    # errors: other species mask, not a checkpoint, truncated file
    p = rcm.default_params()
    p.species_mask = 0x3
    c = rcm.Solver(0, p)
    c.set_repwvl_table_from(rcm.Table(table_path(20)))
    with pytest.raises(rcm.RcmError):
        c.load_checkpoint(path)
    with pytest.raises(rcm.RcmError):
        b.load_checkpoint(table_path(20))
    raw = open(path, "rb").read()
    open(path, "wb").write(raw[: len(raw) // 2])
    with pytest.raises(rcm.RcmError):
        b.load_checkpoint(path)
    with pytest.raises(rcm.RcmError):
        b.load_checkpoint(str(tmp_path / "missing.ckpt"))
    b.close()
    c.close()


def test_lbl_restart_is_bit_identical(rcm, tmp_path):
    atm = rcm.read_atm(os.path.join(GOLDEN, "column21.lbl.atm"))
    full = rcm.read_atm(os.path.join(GOLDEN, "column21.atm"))
    pl = atm[:, 1].copy()
    n, nwvl = 9, 800
    Tlev, vlev = rcm.make_ensemble(n, 4242, pl, atm[:, 2].copy(), full[:, 4:9].T.copy())
    st = rcm.init_columns(pl, Tlev, vlev)
    h2o_ref, o3_ref = st["vmr9"][0, 0].copy(), st["vmr9"][0, 2].copy()
    wvl, tau5 = rcm.make_lbl_tables(nwvl, 777, pl, h2o_ref, o3_ref)

    def fresh():
        s = rcm.Solver(0)
        s.set_lbl_tables(wvl, tau5, h2o_ref, o3_ref, 2.0)
        return s

    a = fresh()
    a.set_columns(pl, st["Tlayer"], Tlev[:, 20].copy(), st["vmr9"], st["rel_hum"])
    a.advance(2)
    path = str(tmp_path / "lbl.ckpt")
    a.save_checkpoint(path)
    a.advance(3)
    end_a = a.get_state()
    a.close()
    b = fresh()
    b.load_checkpoint(path)
    b.advance(3)
    assert same(b.get_state(), end_a)
    r = rcm.Solver(0)  # a repwvl solver refuses an LBL checkpoint
    r.set_repwvl_table_from(rcm.Table(table_path(10)))
    with pytest.raises(rcm.RcmError):
        r.load_checkpoint(path)
    r.close()
    b.close()
