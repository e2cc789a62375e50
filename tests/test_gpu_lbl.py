"""Line-by-line path on the GPU against the CPU restatement (oracle/rcm_oracle.c, rcmo_lbl_*).

The reference has no LBL driver and its tables are not distributed: parity here is against the
builder-defined composition of the reference's components on synthetic tables
written in the reference's text format and read back through the drop-in reader.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return float(np.max(np.abs(a - b) / np.max(np.abs(b), axis=-1, keepdims=True)))


@pytest.fixture(scope="module")
def lbl_case(rcm, tmp_path_factory):
    atm = rcm.read_atm(os.path.join(GOLDEN, "column21.lbl.atm"))       # fpda.lbl.atm: z p T air H2O O3
    full = rcm.read_atm(os.path.join(GOLDEN, "column21.atm"))
    pl = atm[:, 1].copy()
    ncol, nwvl = 24, 1500
    Tlev, vlev = rcm.make_ensemble(ncol, 4242, pl, atm[:, 2].copy(), full[:, 4:9].T.copy())
    st = rcm.init_columns(pl, Tlev, vlev)
    h2o_ref, o3_ref = st["vmr9"][0, 0].copy(), st["vmr9"][0, 2].copy()   # member 0 = the table's atmosphere
    wvl, tau5 = rcm.make_lbl_tables(nwvl, 777, pl, h2o_ref, o3_ref)
    d = tmp_path_factory.mktemp("lbl")
    names = ["h2o", "co2", "o3", "ch4", "n2o"]
    back = []
    for k, nm in enumerate(names):                                        # reference file format round trip
        path = str(d / f"lbl.{nm}.asc")
        rcm.write_lbl_asc(path, wvl, tau5[k])
        stt, x, y = rcm.ascii_file2xy2D(path)
        assert stt == 0 and np.array_equal(x, wvl) and np.array_equal(y, tau5[k])
        back.append(y)
    return dict(pl=pl, st=st, Tsurf=Tlev[:, 20].copy(), wvl=wvl, tau5=np.stack(back), h2o_ref=h2o_ref, o3_ref=o3_ref,
                ncol=ncol)


@pytest.mark.parametrize("co2_factor", [1.0, 2.0])
def test_lbl_steps_match_cpu_restatement(rcm, port, lbl_case, co2_factor):
    c = lbl_case
    solar = rcm.solar_setup()["solar_irr"]
    st = c["st"]
    o3_scale = st["vmr9"][:, 2] / c["o3_ref"]
    s = rcm.Solver(0)
    s.set_lbl_tables(c["wvl"], c["tau5"], c["h2o_ref"], c["o3_ref"], co2_factor)
    s.set_columns(c["pl"], st["Tlayer"], c["Tsurf"], st["vmr9"], st["rel_hum"])
    for nsteps, first in ((1, 0), (2, 1)):
        sc = s.advance(nsteps)
        got = s.get_state()
        ref = port.lbl_advance(c["wvl"], c["tau5"], c["pl"], st["rel_hum"], c["h2o_ref"], o3_scale, co2_factor, solar,
                               st["Tlayer"], c["Tsurf"], st["vmr9"][:, 0], first + nsteps)
        assert relerr(got["E_up"], ref["E_up"]) < 1e-10      # north_star: fluxes within 1e-10 relative
        assert relerr(got["E_down"], ref["E_down"]) < 1e-10
        scale = np.max(np.abs(ref["E_up"]), axis=-1, keepdims=True)
        assert float(np.max(np.abs(got["dE"] - ref["dE"]) / scale)) < 1e-10
        np.testing.assert_allclose(got["Tlayer"], ref["Tlayer"], rtol=1e-10)
        np.testing.assert_allclose(got["Tsurf"], ref["Tsurf"], rtol=1e-10)
        np.testing.assert_allclose(got["h2o"], ref["h2o"], rtol=1e-10)
        np.testing.assert_allclose(got["dt"], ref["dt"], rtol=1e-8)
        assert np.all(np.isfinite(sc))
    s.close()


def test_lbl_doubling_co2_reduces_olr(rcm, lbl_case):
    """2xCO2 forcing (config 5): instantaneous OLR drops, by a few W/m2 for the synthetic CO2 band."""
    c, st = lbl_case, lbl_case["st"]
    olr = {}
    for f in (1.0, 2.0):
        s = rcm.Solver(0)
        s.set_lbl_tables(c["wvl"], c["tau5"], c["h2o_ref"], c["o3_ref"], f)
        s.set_columns(c["pl"], st["Tlayer"], c["Tsurf"], st["vmr9"], st["rel_hum"])
        s.advance(1)
        olr[f] = s.get_state()["E_up"][:, 0]
        s.close()
    assert np.all(olr[2.0] < olr[1.0]) and np.all(olr[1.0] - olr[2.0] < 30.0)


def test_lbl_wavelength_chunking_is_deterministic(rcm, lbl_case):
    """Same columns replicated: every copy gets bit-identical fluxes (fixed-order chunk sums)."""
    c, st = lbl_case, lbl_case["st"]
    rep = 40
    tile = lambda a: np.tile(a, (rep,) + (1,) * (a.ndim - 1))
    s = rcm.Solver(0)
    s.set_lbl_tables(c["wvl"], c["tau5"], c["h2o_ref"], c["o3_ref"], 1.0)
    s.set_columns(c["pl"], tile(st["Tlayer"]), tile(c["Tsurf"]), tile(st["vmr9"]), tile(st["rel_hum"]))
    s.advance(1)
    got = s.get_state()
    n = c["ncol"]
    for k in ("E_up", "E_down", "Tlayer"):
        b = got[k].reshape(rep, n, -1)
        assert np.array_equal(b[0], b[-1]) and np.array_equal(b[0], b[rep // 2])
    s.close()


def test_lbl_shards_are_bit_identical_to_the_whole_ensemble(rcm, lbl_case):
    """The LBL spectral sum runs over fixed 128-wavelength chunks (registers over a chunk's rounds, groups, chunks in
    index order): independent of how many columns a GPU owns.  The ensemble stepped as a whole and as three sequential
    shards gives the same bytes (round 1 sized the chunks from the column count: another summation order per shard)."""
    from our_first_climate_model_b200.distributed import shard_range
    c, st = lbl_case, lbl_case["st"]
    n = c["ncol"]

    def run(lo, hi):
        s = rcm.Solver(0)
        s.set_lbl_tables(c["wvl"], c["tau5"], c["h2o_ref"], c["o3_ref"], 2.0)
        s.set_columns(c["pl"], st["Tlayer"][lo:hi], c["Tsurf"][lo:hi], st["vmr9"][lo:hi], st["rel_hum"][lo:hi])
        s.advance(1)
        s.advance(2)
        out = s.get_state()
        s.close()
        return out

    whole = run(0, n)
    parts = [run(*shard_range(n, r, 3)) for r in range(3)]
    for k in ("E_up", "E_down", "dE", "Tlayer", "Tsurf", "h2o", "dt"):
        assert np.array_equal(np.concatenate([p[k] for p in parts]), whole[k]), k
