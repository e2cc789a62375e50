"""Regression tests for defects found in review (ADVICE.md of round 1)."""
import numpy as np
import pytest

from conftest import table_path

pytestmark = pytest.mark.gpu


def _load(s, rcm, golden, n=20, sl=slice(0, 5)):
    s.set_repwvl_table_from(rcm.Table(table_path(n)))
    s.set_columns(golden["plevel"], golden["Tlayer"][sl], golden["Tsurf"][sl], golden["vmr9"][sl], golden["rel_hum"][sl])


def test_scalars_after_a_longer_call(rcm, golden):
    """advance(8) then advance(3) on the same solver: the ticket counters of the scalar reduction used to sit behind the
    partials of the CURRENT nsteps and aliased partials left by the longer call - no CTA was 'last', the scalars of
    the short call were never written.  Compare with a fresh solver that only ever ran 3 steps at a time."""
    a, b = rcm.Solver(0), rcm.Solver(0)
    try:
        _load(a, rcm, golden)
        a.advance(8)
        sa = a.advance(3)
        sa1 = a.advance(1)
        _load(b, rcm, golden)
        for _ in range(2):
            b.advance(3)
        b.advance(2)
        sb = b.advance(3)
        sb1 = b.advance(1)
        assert np.array_equal(sa, sb) and np.array_equal(sa1, sb1)
        assert np.all(sa[:, 3] > 0)  # max|dE| of a real step is never 0
        # ... and the C driver with a short last block (max_steps % check_every != 0) decides on fresh scalars
        _load(a, rcm, golden)
        done, last = a.run_to_equilibrium(11, 8)
        _load(b, rcm, golden)
        ref = b.advance(11)
        assert done == 11 and np.array_equal(last, ref[-1])
    finally:
        a.close()
        b.close()


def test_lbl_tables_invalidate_the_repwvl_grid(rcm, golden):
    """rcm_set_lbl_tables re-sizes the wavelength axis: the repwvl-only entry points must refuse to run on the device
    arrays of the old grid (they used to launch over LBL nwvl with buffers sized for the repwvl table)."""
    s = rcm.Solver(0)
    try:
        _load(s, rcm, golden)
        s.build_tau()
        atm_h2o = golden["vmr9"][0, 0]
        wvl, tau5 = rcm.make_lbl_tables(300, 5, golden["plevel"], atm_h2o, golden["vmr9"][0, 2])
        s.set_lbl_tables(wvl, tau5, atm_h2o, None, 1.0)
        with pytest.raises(rcm.RcmError):
            s.build_tau()
        with pytest.raises(rcm.RcmError):
            s.radiative_transfer(np.zeros((5, 300, 20)))
        s.advance(1)  # the LBL step itself works
        _load(s, rcm, golden)  # and a repwvl table brings the repwvl path back
        tau, _, _ = s.build_tau()
        assert np.array_equal(tau, golden["tau20"][:5])
    finally:
        s.close()


def test_checkpoint_refuses_another_wavelength_count(rcm, golden, tmp_path):
    s = rcm.Solver(0)
    try:
        _load(s, rcm, golden, 20)
        s.advance(2)
        path = str(tmp_path / "c.ckpt")
        s.save_checkpoint(path)
        s.set_repwvl_table_from(rcm.Table(table_path(10)))
        with pytest.raises(rcm.RcmError):
            s.load_checkpoint(path)
        s.set_repwvl_table_from(rcm.Table(table_path(20)))
        s.load_checkpoint(path)
        s.advance(1)
    finally:
        s.close()


def test_planck_exponent_stays_in_range(rcm, golden):
    """The kernels' exp has no overflow path (robustness): a runaway cold column (2 K) must give finite - zero -
    thermal emission like the reference's exp -> inf -> B = 0, not a wrapped exponent; a spectral grid that would put the
    Planck exponent out of range at atmospheric temperatures (wavelengths below ~1 um) is refused at the ABI."""
    s = rcm.Solver(0)
    try:
        s.set_repwvl_table_from(rcm.Table(table_path(20)))
        T = golden["Tlayer"][:3].copy()
        T[1, :] = 2.0
        T[2, 5] = 0.5
        Ts = np.array([288.2, 2.0, 288.2])
        for path in (0, 1):
            s.set_option(5, path)
            s.set_columns(golden["plevel"], T, Ts, golden["vmr9"][:3], golden["rel_hum"][:3])
            tau, _, _ = s.build_tau()
            s.update_columns(Tlayer=T, Tsurf=Ts)
            Ed, Eu, dE = s.radiative_transfer(np.clip(tau, 0, None))
            assert np.isfinite(Ed).all() and np.isfinite(Eu).all() and np.isfinite(dE).all()
            # nothing radiates at 2 K; the source is evaluated at T_floor = 5.1 K: below a 5 K blackbody's 3.8e-5 W/m2
            assert np.all(np.abs(Eu[1]) < 1e-5) and np.all(np.abs(Ed[1]) < 1e-5)
            assert Eu[2, 0] > 50.0 and Eu[0, 0] > 50.0
            s.set_columns(golden["plevel"], T, Ts, golden["vmr9"][:3], golden["rel_hum"][:3])
            s.advance(1)
            st = s.get_state()
            # (in the full step the 2 K column's cross sections are extrapolated 100 K below the table and come out negative:
            # its fluxes are meaningless - amplified by exp(+8/mu) per layer, inf / NaN in the reference too; its neighbours in
            # the tile are untouched)
            assert np.isfinite(st["E_up"][[0, 2]]).all() and np.isfinite(st["Tlayer"][[0, 2]]).all()
            assert abs(st["E_up"][0, 0] - golden["s1_E_up_20"][0, 0]) < 1e-10 * st["E_up"][0, 0]
        s.set_option(5, 0)
        with pytest.raises(rcm.RcmError, match="1043"):
            s.set_spectral_grid(np.array([500.0, 5000.0]), np.array([1.0, 1.0]))
        s.set_spectral_grid(np.array([1100.0, 5000.0]), np.array([1.0, 1.0]))
    finally:
        s.close()


def test_abi_refuses_bad_calls_without_touching_the_device(rcm, golden):
    """Error behaviour of the C ABI: status codes, never exit / throw; a refused call leaves the
    solver usable.  The reference's entry points have no such checks ('size mismatch only prints')."""
    import ctypes as C
    L = rcm.load_library()
    ARG, STATE = 1, 2
    assert L.rcm_status_string(ARG) and L.rcm_status_string(STATE)
    s = rcm.Solver(0)
    try:
        h = s._h
        sc = (rcm.StepScalars * 4)()
        # nothing loaded yet
        assert L.rcm_advance(h, 1, sc) == STATE and b"table and columns" in L.rcm_last_error(h)
        assert L.rcm_get_state(h, None, None, None, None, None, None, None, None) == STATE
        assert L.rcm_step_host(h, None, None, None, None, None, None, None, None) != 0
        assert L.rcm_save_checkpoint(h, b"/tmp/never_written.ckpt") != 0
        # NULL handle / NULL or empty inputs
        assert L.rcm_advance(None, 1, sc) == ARG and L.rcm_column_count(None) <= 0
        d = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.c_void_p)
        pl, T, Ts, v, rh = (golden[k] for k in ("plevel", "Tlayer", "Tsurf", "vmr9", "rel_hum"))
        assert L.rcm_set_columns(h, 0, d(pl), d(T), d(Ts), d(v), d(rh)) == ARG
        assert L.rcm_set_columns(h, -3, d(pl), d(T), d(Ts), d(v), d(rh)) == ARG
        assert L.rcm_set_columns(h, 4, d(pl), None, d(Ts), d(v), d(rh)) == ARG
        # columns without a table: still a state error, then the normal sequence works on the same handle
        s.set_columns(pl, T[:4], Ts[:4], v[:4], rh[:4])
        assert L.rcm_advance(h, 1, sc) == STATE
        s.set_repwvl_table_from(rcm.Table(table_path(20)))
        assert L.rcm_advance(h, 0, sc) == ARG and L.rcm_advance(h, -1, sc) == ARG
        assert L.rcm_advance(h, 1, sc) == 0
        st = s.get_state()
        assert abs(st["E_up"][0, 0] - golden["s1_E_up_20"][0, 0]) < 1e-10 * st["E_up"][0, 0]
        assert s.launch_count() > 0
    finally:
        s.close()
