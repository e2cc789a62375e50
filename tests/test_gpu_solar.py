"""Per-column solar setup on the device (doubling_adding + solar_radiative_transfer_setup,
main.cpp:214-264, batched) and its use by the step: per-column solar_irr in the heating of the lowest layer
(main.cpp:341) and per-column grey-cloud optical depth in tau (main.cpp:266-274).

Checkers: the oracle port (oracle/rcm_oracle.c): its solar setup, which reproduces the solar lines of the reference's
committed output.txt (tests/test_oracle.py), and its time stepping.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, table_path

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.max(np.abs(a - b) / np.max(np.abs(b), axis=-1, keepdims=True)))


def ensemble(rcm, n, seed):
    atm = rcm.read_atm(os.path.join(GOLDEN, "column21.atm"))
    pl = atm[:, 1].copy()
    Tlev, vlev = rcm.make_ensemble(n, seed, pl, atm[:, 2].copy(), atm[:, 4:9].T.copy())
    st = rcm.init_columns(pl, Tlev, vlev)
    return pl, st, Tlev[:, 20].copy()


def forcing(n, seed, wide=False):
    """Per-column cloud optical depth, zenith cosine, surface albedo.  wide: the whole plausible range (solar setup
    alone); otherwise a range in which every column keeps a positive max(dE) - the reference's time step divides by the
    signed maximum (main.cpp:157) and turns a column without net heating anywhere into NaNs."""
    rng = np.random.default_rng(seed)
    if wide:
        return rng.uniform(0.2, 6.0, n), rng.uniform(0.1, 1.0, n), rng.uniform(0.0, 0.9, n)
    return rng.uniform(0.5, 4.0, n), rng.uniform(0.3, 0.8, n), rng.uniform(0.05, 0.3, n)


def test_device_solar_setup_equals_oracle(rcm, port):
    """4,099 columns (ragged last block) with their own cloud optical depth, zenith cosine and albedo."""
    n = 4099
    pl, st, Ts = ensemble(rcm, n, 7)
    tau_s, mu_s, alb = forcing(n, 11, wide=True)
    s = rcm.Solver(0)
    s.set_repwvl_table_from(rcm.Table(table_path(10)))
    s.set_columns(pl, st["Tlayer"], Ts, st["vmr9"], st["rel_hum"])
    got = s.set_column_solar(None, tau_s, mu_s, alb)
    ref_irr, ref_rt = np.zeros(n), np.zeros(n)
    for c in range(n):  # the checker is the oracle's restatement of main.cpp:214-264 (oracle/rcm_oracle.c)
        o = port.solar_setup(tau_s=tau_s[c], mu_s=mu_s[c], albedo=alb[c])
        ref_irr[c], ref_rt[c] = o["solar_irr"], o["r_total"]
    # same operation order, no contraction; the only library call is pow(t_dir, 2) on the host (RN square here)
    np.testing.assert_allclose(got["r_total"], ref_rt, rtol=4e-16, atol=0)
    np.testing.assert_allclose(got["solar_irr"], ref_irr, rtol=4e-16, atol=0)
    assert np.mean(got["solar_irr"] == ref_irr) > 0.95
    # scalars only = the reference's committed Consts: bit-identical to the value pinned by output.txt
    one = s.set_column_solar(rcm.default_solar_params())
    assert np.all(one["solar_irr"] == port.solar_setup()["solar_irr"]) and np.all(one["r_total"] == port.solar_setup()["r_total"])
    assert f"{one['solar_irr'][0]:.6f}" == "236.882897"  # tau_s = 2.0, the committed Consts (tests/test_oracle.py)
    s.close()


@pytest.mark.parametrize("cloud_from_tau_s", [False, True])
def test_step_with_per_column_solar_and_cloud_matches_port(rcm, port, cloud_from_tau_s):
    n = 21  # ragged: one full 16-column tile + 5
    pl, st, Ts = ensemble(rcm, n, 99)
    tau_s, mu_s, alb = forcing(n, 5)
    tab = port.load_rcmtab(table_path(20))
    s = rcm.Solver(0)
    s.set_repwvl_table_from(rcm.Table(table_path(20)))
    s.set_columns(pl, st["Tlayer"], Ts, st["vmr9"], st["rel_hum"])
    f = s.set_column_solar(None, tau_s, mu_s, alb, cloud_from_tau_s=cloud_from_tau_s)
    sc = s.advance(3)
    got = s.get_state()
    toa = 0.0
    for c in range(n):
        ref = port.advance(tab, pl, st["rel_hum"][c], f["solar_irr"][c], st["Tlayer"][c], Ts[c], st["vmr9"][c], 3,
                           tau_s=tau_s[c] if cloud_from_tau_s else 2.0)
        assert np.all(np.isfinite(ref["Tlayer"])) and np.all(np.isfinite(got["Tlayer"][c]))
        assert relerr(got["E_up"][c:c + 1], ref["E_up"]) < 1e-10
        assert relerr(got["E_down"][c:c + 1], ref["E_down"]) < 1e-10
        np.testing.assert_allclose(got["Tlayer"][c], ref["Tlayer"][0], rtol=1e-10)
        np.testing.assert_allclose(got["Tsurf"][c], ref["Tsurf"][0], rtol=1e-10)
        np.testing.assert_allclose(got["dE"][c], ref["dE"][0], rtol=0, atol=1e-9 * np.abs(ref["dE"]).max())
        toa += f["solar_irr"][c] - ref["E_up"][0, 0]
    assert abs(sc[-1, 0] - toa) < 1e-9 * n * 300  # the TOA diagnostic uses the column's own solar_irr
    # tau alone: the column's cloud sits in the cloud layer, bit-exact
    tau, _, _ = s.build_tau()
    for c in (0, 17, 20):
        tref, _, _ = port.read_tau(tab, pl, got["Tlayer"][c], np.concatenate([got["h2o"][c:c + 1], st["vmr9"][c, 1:]]),
                                   cloud_on=True, tau_s=tau_s[c] if cloud_from_tau_s else 2.0)
        assert np.array_equal(tau[c], tref)
    # back to the ensemble-wide constants
    s.set_columns(pl, st["Tlayer"], Ts, st["vmr9"], st["rel_hum"])
    s.advance(1)
    base = s.get_state()
    s.set_columns(pl, st["Tlayer"], Ts, st["vmr9"], st["rel_hum"])
    s.set_column_solar(None, tau_s, mu_s, alb, cloud_from_tau_s=cloud_from_tau_s)
    s.set_column_solar(clear=True)
    s.advance(1)
    again = s.get_state()
    assert np.array_equal(base["Tlayer"], again["Tlayer"]) and np.array_equal(base["dE"], again["dE"])
    s.close()


def test_lbl_step_with_per_column_solar_and_cloud(rcm, port):
    atm = rcm.read_atm(os.path.join(GOLDEN, "column21.lbl.atm"))
    full = rcm.read_atm(os.path.join(GOLDEN, "column21.atm"))
    pl = atm[:, 1].copy()
    n, nwvl = 6, 600
    Tlev, vlev = rcm.make_ensemble(n, 4242, pl, atm[:, 2].copy(), full[:, 4:9].T.copy())
    st = rcm.init_columns(pl, Tlev, vlev)
    h2o_ref, o3_ref = st["vmr9"][0, 0].copy(), st["vmr9"][0, 2].copy()
    wvl, tau5 = rcm.make_lbl_tables(nwvl, 777, pl, h2o_ref, o3_ref)
    tau_s, mu_s, alb = forcing(n, 3)
    s = rcm.Solver(0)
    s.set_lbl_tables(wvl, tau5, h2o_ref, o3_ref, 2.0)
    s.set_columns(pl, st["Tlayer"], Tlev[:, 20].copy(), st["vmr9"], st["rel_hum"])
    f = s.set_column_solar(None, tau_s, mu_s, alb, cloud_from_tau_s=True)
    s.advance(2)
    got = s.get_state()
    for c in range(n):
        ref = port.lbl_advance(wvl, tau5, pl, st["rel_hum"][c], h2o_ref, st["vmr9"][c, 2] / o3_ref, 2.0,
                               f["solar_irr"][c], st["Tlayer"][c], Tlev[c, 20], st["vmr9"][c, 0], 2, tau_s=tau_s[c])
        assert relerr(got["E_up"][c:c + 1], ref["E_up"]) < 1e-9
        np.testing.assert_allclose(got["Tlayer"][c], ref["Tlayer"][0], rtol=1e-10)
    s.close()


def test_column_solar_argument_errors(rcm):
    s = rcm.Solver(0)
    with pytest.raises(Exception):
        s.set_column_solar(None, 1.0)  # no columns yet
    pl, st, Ts = ensemble(rcm, 4, 1)
    p = rcm.default_params()
    p.cloud_layer = -1
    s2 = rcm.Solver(0, p)
    s2.set_repwvl_table_from(rcm.Table(table_path(10)))
    s2.set_columns(pl, st["Tlayer"], Ts, st["vmr9"], st["rel_hum"])
    with pytest.raises(Exception):
        s2.set_column_solar(None, 1.0, cloud_from_tau_s=True)  # no cloud layer to put it in
    sp = rcm.default_solar_params()
    sp.doublings = 99
    with pytest.raises(Exception):
        s2.set_column_solar(sp)
    s.close()
    s2.close()


def test_column_solar_after_the_ensemble_grew(rcm, port):
    """Capacity regression: forcing buffers allocated for a small ensemble must follow a later, bigger rcm_set_columns."""
    s = rcm.Solver(0)
    s.set_repwvl_table_from(rcm.Table(table_path(10)))
    pl, st, Ts = ensemble(rcm, 8, 3)
    s.set_columns(pl, st["Tlayer"], Ts, st["vmr9"], st["rel_hum"])
    s.set_column_solar(None, *forcing(8, 1))
    n = 5000
    pl, st, Ts = ensemble(rcm, n, 4)
    s.set_columns(pl, st["Tlayer"], Ts, st["vmr9"], st["rel_hum"])
    tau_s, mu_s, alb = forcing(n, 2)
    f = s.set_column_solar(None, tau_s, mu_s, alb, cloud_from_tau_s=True)
    s.advance(1)
    got = s.get_state()
    tab = port.load_rcmtab(table_path(10))
    for c in (0, 4999):
        ref = port.advance(tab, pl, st["rel_hum"][c], f["solar_irr"][c], st["Tlayer"][c], Ts[c], st["vmr9"][c], 1, tau_s=tau_s[c])
        assert relerr(got["E_up"][c:c + 1], ref["E_up"]) < 1e-10
        np.testing.assert_allclose(got["dE"][c], ref["dE"][0], rtol=0, atol=1e-9 * np.abs(ref["dE"]).max())
    s.close()
