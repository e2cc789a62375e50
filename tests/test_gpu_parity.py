"""Parity of the CUDA path (through the C ABI) against the reference.

Checker = committed golden vectors produced by the UNMODIFIED reference (tools/make_golden.py),
plus the plain-C oracle port on seeded inputs.  Tolerances: table indices and tau bit-exact;
FP64 fluxes / heating rates 1e-10 relative (north_star) after one step; equilibrium 1e-3 K.
"""
import numpy as np
import pytest

from conftest import table_path

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def relerr(a, b, scale=None):
    """max |a-b| / scale, scale = per-column max |b| (fluxes of one column share a magnitude)."""
    a, b = np.asarray(a), np.asarray(b)
    if scale is None:
        scale = np.max(np.abs(b), axis=-1, keepdims=True)
    return float(np.max(np.abs(a - b) / scale))


@pytest.fixture(scope="module")
def solver(rcm):
    s = rcm.Solver(0)
    yield s
    s.close()


def load(solver, rcm, golden, n, sl=slice(None)):
    solver.set_repwvl_table_from(rcm.Table(table_path(n)))
    solver.set_columns(golden["plevel"], golden["Tlayer"][sl], golden["Tsurf"][sl], golden["vmr9"][sl],
                       golden["rel_hum"][sl])


@pytest.mark.parametrize("n", [10, 20, 100])
def test_tau_and_indices_bit_exact(solver, rcm, golden, n):
    load(solver, rcm, golden, n)
    tau, lp, lt = solver.build_tau()
    assert np.array_equal(lp, golden[f"lowpos_p{n}"])
    assert np.array_equal(lt, golden[f"lowpos_t{n}"])
    assert np.array_equal(tau, golden[f"tau{n}"]), "tau must be bit-identical to read_tau + cloud_into_tau"


@pytest.mark.parametrize("n", [10, 100])
def test_tau_edge_members_bit_exact(solver, rcm, golden_edge, n):
    e = golden_edge
    solver.set_repwvl_table_from(rcm.Table(table_path(n)))
    solver.set_columns(e["plevel"], e["Tlayer"], np.full(4, 288.2), e["vmr9"], np.zeros((4, 20)))
    tau, lp, lt = solver.build_tau()
    assert np.array_equal(lt, e[f"lowpos_t{n}"])          # out-of-range -> last interval, exact hit -> lower
    assert np.array_equal(lp[0], e["lowpos_p"])
    assert np.array_equal(tau, e[f"tau{n}"])


@pytest.mark.parametrize("n", [10, 20, 100])
@pytest.mark.parametrize("cubes", [1, 0])
def test_radiative_transfer_given_tau(solver, rcm, golden, n, cubes):
    """K2-K4 alone: feed the reference's tau, compare E_down, E_up, dE of step 0."""
    solver.set_option(0, cubes)
    load(solver, rcm, golden, n)
    # step 0 of the reference sorts theta before the radiative transfer: sorted T is in the trace-free
    # golden as the input of s1 only implicitly, so sort on the host exactly as main.cpp:536-540
    T = golden["Tlayer"] * golden["conv"]
    T = -np.sort(-T, axis=1) / golden["conv"]
    solver.update_columns(Tlayer=T)
    Ed, Eu, dE = solver.radiative_transfer(golden[f"tau{n}"])
    solver.set_option(0, 1)
    assert relerr(Ed, golden[f"s1_E_down_{n}"]) < RTOL
    assert relerr(Eu, golden[f"s1_E_up_{n}"]) < RTOL
    # heating rates are differences of fluxes: tolerance relative to the column's flux scale
    scale = np.max(np.abs(golden[f"s1_E_up_{n}"]), axis=-1, keepdims=True)
    assert relerr(dE, golden[f"s1_dE_{n}"], scale) < RTOL
    assert relerr(dE, golden[f"s1_dE_{n}"]) < RTOL  # and relative to max|dE| of the column


@pytest.mark.parametrize("nw", [7, 150])
def test_radiative_transfer_on_a_bare_spectral_grid(rcm, port, golden, nw):
    """rcm_set_spectral_grid + rcm_radiative_transfer (what the `radiative_transfer` adapter does) with a tau that was built
    elsewhere: 7 wavelengths, and 150 - more than the step kernel keeps Planck factors for in shared memory (PLK_MAX)."""
    rng = np.random.default_rng(nw)
    ncol = 13  # ragged: less than one 16-column tile
    tab = port.load_rcmtab(table_path(100))
    pick = rng.integers(0, 100, nw)
    wvl = tab["wvl"][pick] * rng.uniform(0.98, 1.02, nw)
    weight = tab["weight"][pick] * rng.uniform(0.5, 1.5, nw)
    tau = golden["tau100"][:ncol][:, pick, :] * rng.uniform(0.5, 2.0, (ncol, nw, 1))
    T = golden["Tlayer"][:ncol] + rng.uniform(-3, 3, (ncol, 20))
    Ts = golden["Tsurf"][:ncol] + rng.uniform(-2, 2, ncol)
    s = rcm.Solver(0)
    s.set_spectral_grid(wvl, weight)
    s.set_columns(golden["plevel"], T, Ts, golden["vmr9"][:ncol], golden["rel_hum"][:ncol])
    Ed, Eu, dE = s.radiative_transfer(tau)
    for c in range(ncol):
        rEd, rEu, rdE = port.radiative_transfer(tau[c], wvl, weight, T[c], Ts[c], float(golden["solar_irr"]))
        scale = np.abs(rEu).max()
        assert np.max(np.abs(Ed[c] - rEd)) < RTOL * scale and np.max(np.abs(Eu[c] - rEu)) < RTOL * scale
        assert np.max(np.abs(dE[c] - rdE)) < RTOL * scale
    s.close()


@pytest.mark.parametrize("n", [20, 100])
def test_angle_pair_units_on_and_off(solver, rcm, golden, n):
    """RCM_OPT_ANGLE_PAIRS: chain heads that share a virtual root take it from one exp (default) - against the golden
    fluxes with and without, and against each other (same mu values, other summation order and a few more roundings)."""
    out = {}
    for pairs in (1, 0):
        solver.set_option(4, pairs)
        load(solver, rcm, golden, n)
        solver.advance(1)
        st = solver.get_state()
        assert relerr(st["E_down"], golden[f"s1_E_down_{n}"]) < RTOL
        assert relerr(st["E_up"], golden[f"s1_E_up_{n}"]) < RTOL
        out[pairs] = st
    solver.set_option(4, 1)
    assert relerr(out[1]["E_up"], out[0]["E_up"]) < 1e-13 and relerr(out[1]["E_down"], out[0]["E_down"]) < 1e-13
    assert not np.array_equal(out[1]["E_up"], out[0]["E_up"])  # the option really changes the schedule


@pytest.mark.parametrize("n", [10, 20, 100])
def test_one_fused_step(solver, rcm, golden, n):
    load(solver, rcm, golden, n)
    sc = solver.advance(1)
    st = solver.get_state()
    assert relerr(st["E_down"], golden[f"s1_E_down_{n}"]) < RTOL
    assert relerr(st["E_up"], golden[f"s1_E_up_{n}"]) < RTOL
    scale = np.max(np.abs(golden[f"s1_E_up_{n}"]), axis=-1, keepdims=True)
    assert relerr(st["dE"], golden[f"s1_dE_{n}"], scale) < RTOL
    assert relerr(st["dE"], golden[f"s1_dE_{n}"]) < RTOL  # heating rates: 1e-10 also relative to the column's own max|dE|
    np.testing.assert_allclose(st["dt"], golden[f"s1_dt_{n}"], rtol=1e-10)
    np.testing.assert_allclose(st["Tlayer"], golden[f"s1_Tlayer_{n}"], rtol=1e-11)
    np.testing.assert_allclose(st["Tsurf"], golden[f"s1_Tsurf_{n}"], rtol=1e-11)
    np.testing.assert_allclose(st["time_h"], golden[f"s1_time_h_{n}"], rtol=1e-6)
    toa = golden["solar_irr"] - golden[f"s1_E_up_{n}"][:, 0]
    np.testing.assert_allclose(sc[0, 0], toa.sum(), rtol=1e-10)
    np.testing.assert_allclose(sc[0, 3], np.abs(golden[f"s1_dE_{n}"]).max(), rtol=1e-10)


@pytest.mark.parametrize("n", [20, 100])
@pytest.mark.parametrize("chunks", [(5,), (1, 1, 3), (2, 3)])
def test_five_steps_fused_and_chunked(solver, rcm, golden, n, chunks):
    """5 reference iterations, as one launch and split over several launches: same trajectory."""
    load(solver, rcm, golden, n)
    for k in chunks:
        solver.advance(k)
    st = solver.get_state()
    np.testing.assert_allclose(st["Tlayer"], golden[f"s5_Tlayer_{n}"], rtol=1e-10)
    np.testing.assert_allclose(st["Tsurf"], golden[f"s5_Tsurf_{n}"], rtol=1e-10)
    np.testing.assert_allclose(st["h2o"], golden[f"s5_h2o_{n}"], rtol=1e-10)
    assert relerr(st["E_up"], golden[f"s5_E_up_{n}"]) < 1e-9
    assert relerr(st["E_down"], golden[f"s5_E_down_{n}"]) < 1e-9
    np.testing.assert_allclose(st["dt"], golden[f"s5_dt_{n}"], rtol=1e-8)
    np.testing.assert_allclose(st["time_h"], golden[f"s5_time_h_{n}"], rtol=1e-6)


def test_300_steps_profile(solver, rcm, golden):
    load(solver, rcm, golden, 100, slice(0, 4))
    solver.advance(300, want_scalars=False)
    st = solver.get_state()
    assert np.max(np.abs(st["Tlayer"] - golden["s300_Tlayer_100"])) < 1e-6
    assert np.max(np.abs(st["Tsurf"] - golden["s300_Tsurf_100"])) < 1e-6


def test_equilibrium_profile_within_1e_3_K(solver, rcm, golden):
    """north_star: the equilibrium temperature profile within 1e-3 K (6000 reference iterations)."""
    load(solver, rcm, golden, 100, slice(0, 1))
    sc = None
    for _ in range(12):
        sc = solver.advance(500)
    st = solver.get_state()
    assert np.max(np.abs(st["Tlayer"] - golden["s6000_Tlayer_100"])) < 1e-3
    assert abs(st["Tsurf"][0] - golden["s6000_Tsurf_100"][0]) < 1e-3
    assert sc[-1, 1] < 1e-2  # stationarity diagnostic: the sorted profile has stopped moving


def test_run_to_equilibrium_driver(solver, rcm, golden):
    """The RCE driver loop: with an unreachable threshold it runs exactly max_steps (= the golden 6000-iteration
    profile); with 1e-2 K per step it stops early with every column converged (the approach is slow: a few K away)."""
    load(solver, rcm, golden, 100, slice(0, 1))
    p = rcm.default_params()
    p.solar_irr = float(golden["solar_irr"])
    p.dT_converged = 0.0
    solver.set_params(p)
    done, last = solver.run_to_equilibrium(6000, 750)
    st = solver.get_state()
    assert done == 6000 and last[2] == 0
    assert np.max(np.abs(st["Tlayer"] - golden["s6000_Tlayer_100"])) < 1e-3
    p.dT_converged = 1e-2
    solver.set_params(p)
    load(solver, rcm, golden, 100, slice(0, 3))
    done, last = solver.run_to_equilibrium(6000, 250)
    assert done < 6000 and done % 250 == 0 and last[2] == 3 and last[1] < 1e-2
    st = solver.get_state()
    assert np.max(np.abs(st["Tlayer"][0] - golden["s6000_Tlayer_100"][0])) < 5.0
    p.dT_converged = rcm.default_params().dT_converged
    solver.set_params(p)


def test_distributed_driver_single_rank_equals_c_driver(solver, rcm, golden):
    """distributed.run_to_equilibrium (the N-rank loop; here one rank, no process group) == rcm_run_to_equilibrium."""
    import torch
    from our_first_climate_model_b200 import distributed as rdist
    p = rcm.default_params()
    p.solar_irr = float(golden["solar_irr"])
    p.dT_converged = 5e-2
    solver.set_params(p)
    load(solver, rcm, golden, 20, slice(0, 5))
    done_c, last_c = solver.run_to_equilibrium(3000, 200)
    T_c = solver.get_state()["Tlayer"]
    load(solver, rcm, golden, 20, slice(0, 5))
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        solver.set_stream(stream.cuda_stream)   # the scalars are read on the stream the solver launches on
        res = rdist.run_to_equilibrium(solver, 5, 3000, 200)
        stream.synchronize()
        solver.set_stream(None)
    assert res["steps"] == done_c and res["converged_fraction"] == 1.0 and res["max_dT"] == last_c[1]
    assert np.array_equal(solver.get_state()["Tlayer"], T_c)
    p.dT_converged = rcm.default_params().dT_converged
    solver.set_params(p)


def test_port_oracle_on_seeded_ensemble(solver, rcm, port, golden):
    """Seeded 200-column ensemble, one step, CUDA vs the plain-C oracle port."""
    atm = rcm.read_atm(table_path(100).replace("Reduced100Forcing.rcmtab", "column21.atm"))
    pl = atm[:, 1]
    Tlev, vlev = rcm.make_ensemble(200, 2024, pl, atm[:, 2], atm[:, 4:9].T.copy())
    st0 = rcm.init_columns(pl, Tlev, vlev)
    tab = port.load_rcmtab(table_path(100))
    ref = port.advance(tab, pl, st0["rel_hum"], golden["solar_irr"], st0["Tlayer"], Tlev[:, 20], st0["vmr9"], 2)
    solver.set_repwvl_table_from(rcm.Table(table_path(100)))
    solver.set_columns(pl, st0["Tlayer"], Tlev[:, 20], st0["vmr9"], st0["rel_hum"])
    solver.advance(2)
    st = solver.get_state()
    assert relerr(st["E_up"], ref["E_up"]) < 1e-9
    assert relerr(st["E_down"], ref["E_down"]) < 1e-9
    np.testing.assert_allclose(st["Tlayer"], ref["Tlayer"], rtol=1e-10)


@pytest.mark.parametrize("nangle,cubes,cloud", [(1, 1, 17), (7, 1, 17), (12, 0, -1), (45, 1, 3), (64, 1, 17), (30, 0, 17)])
def test_other_quadratures_and_clouds_against_the_port(solver, rcm, port, golden, nangle, cubes, cloud):
    """Consts the reference hard-codes, varied: number of angles (odd counts pad the chain schedule, without cubes
    the exponent clamp moves into exp), cloud layer / no cloud; ragged ensemble (37 columns); three steps."""
    atm = rcm.read_atm(table_path(20).replace("Reduced20Forcing.rcmtab", "column21.atm"))
    pl = atm[:, 1]
    Tlev, vlev = rcm.make_ensemble(37, 31 + nangle, pl, atm[:, 2], atm[:, 4:9].T.copy())
    st0 = rcm.init_columns(pl, Tlev, vlev)
    p = rcm.default_params()
    p.solar_irr = float(golden["solar_irr"])
    p.nangle, p.cloud_layer = nangle, cloud
    solver.set_option(0, cubes)
    solver.set_params(p)
    try:
        solver.set_repwvl_table_from(rcm.Table(table_path(20)))
        solver.set_columns(pl, st0["Tlayer"], Tlev[:, 20], st0["vmr9"], st0["rel_hum"])
        solver.advance(3)
        st = solver.get_state()
    finally:
        solver.set_option(0, 1)
        solver.set_params(rcm.default_params())
    ref = port.advance(port.load_rcmtab(table_path(20)), pl, st0["rel_hum"], float(golden["solar_irr"]), st0["Tlayer"],
                       Tlev[:, 20], st0["vmr9"], 3, nangle=nangle, cloud_layer=cloud)
    assert relerr(st["E_up"], ref["E_up"]) < 1e-9
    assert relerr(st["E_down"][:, 1:], ref["E_down"][:, 1:]) < 1e-9
    np.testing.assert_allclose(st["Tlayer"], ref["Tlayer"], rtol=1e-10)
    np.testing.assert_allclose(st["dt"], ref["dt"], rtol=1e-8)


def test_other_species_masks_and_extreme_profiles(solver, rcm, port, golden):
    """(1) species_mask other than the default five (generic-species kernel instantiation, no row staging): only H2O and
    CO2 absorb, the other VMRs are zero as the mask requires.  (2) Far-from-standard profiles: +-60 K offsets with
    vertical structure, VMRs scaled by 10^U(-1,1), no ozone in some columns - table lookups run out of range
    (LowerPos semantics) and tau spans many decades."""
    atm = rcm.read_atm(table_path(20).replace("Reduced20Forcing.rcmtab", "column21.atm"))
    pl = atm[:, 1]
    rng = np.random.default_rng(11)
    ncol = 45
    Tlev = atm[:, 2][None, :] + rng.uniform(-60, 60, (ncol, 1)) + 15 * np.sin(np.linspace(0, 3, 21))[None, :] * rng.uniform(-1, 1, (ncol, 1))
    vlev = np.tile(atm[:, 4:9].T[None], (ncol, 1, 1)) * 10 ** rng.uniform(-1, 1, (ncol, 5, 1))
    vlev[::7, 1] = 0.0
    st0 = rcm.init_columns(pl, Tlev, vlev)
    tab = port.load_rcmtab(table_path(20))
    solar = float(golden["solar_irr"])
    for mask in (0x2F, 0x03):
        vm = st0["vmr9"].copy()
        for k in range(9):
            if not mask >> k & 1:
                vm[:, k] = 0.0
        p = rcm.default_params()
        p.solar_irr, p.species_mask = solar, mask
        solver.set_params(p)
        try:
            solver.set_repwvl_table_from(rcm.Table(table_path(20)))
            solver.set_columns(pl, st0["Tlayer"], Tlev[:, 20], vm, st0["rel_hum"])
            solver.advance(2)
            st = solver.get_state()
        finally:
            solver.set_params(rcm.default_params())
        ref = port.advance(tab, pl, st0["rel_hum"], solar, st0["Tlayer"], Tlev[:, 20], vm, 2)
        assert relerr(st["E_up"], ref["E_up"]) < 1e-9, hex(mask)
        assert relerr(st["E_down"][:, 1:], ref["E_down"][:, 1:]) < 1e-9, hex(mask)
        np.testing.assert_allclose(st["Tlayer"], ref["Tlayer"], rtol=1e-10)
        np.testing.assert_allclose(st["h2o"], ref["vmr9"][:, 0], rtol=1e-10)


def test_row_staging_fallback_and_equivalence(solver, rcm, port, golden):
    """K1 normally reads table rows staged in shared memory (three candidate temperature intervals per layer and
    tile).  (1) Staged and unstaged K1 give bit-identical states.  (2) A tile whose columns span more than three
    intervals takes the global-memory path: checked against the port."""
    atm = rcm.read_atm(table_path(100).replace("Reduced100Forcing.rcmtab", "column21.atm"))
    pl = atm[:, 1]
    ncol = 50
    Tlev, vlev = rcm.make_ensemble(ncol, 5, pl, atm[:, 2], atm[:, 4:9].T.copy())
    wild = Tlev + np.linspace(-55.0, 55.0, ncol)[:, None]          # 110 K across every 16-column tile
    solver.set_repwvl_table_from(rcm.Table(table_path(100)))
    res = {}
    for name, T in (("mild", Tlev), ("wild", wild)):
        st0 = rcm.init_columns(pl, T, vlev)
        for staged in (1, 0):
            solver.set_option(3, staged)
            solver.set_columns(pl, st0["Tlayer"], T[:, 20], st0["vmr9"], st0["rel_hum"])
            solver.advance(2)
            res[name, staged] = solver.get_state()
        solver.set_option(3, 1)
        for k in ("E_up", "E_down", "Tlayer", "dt"):
            assert np.array_equal(res[name, 1][k], res[name, 0][k]), (name, k)
        ref = port.advance(port.load_rcmtab(table_path(100)), pl, st0["rel_hum"], float(golden["solar_irr"]),
                           st0["Tlayer"], T[:, 20], st0["vmr9"], 2)
        assert relerr(res[name, 1]["E_up"], ref["E_up"]) < 1e-9
        np.testing.assert_allclose(res[name, 1]["Tlayer"], ref["Tlayer"], rtol=1e-10)
    tau, lp, lt = solver.build_tau()                               # of the wild ensemble after two steps
    lt = lt[:48].reshape(3, 16, 20)
    assert (lt.max(axis=1) - lt.min(axis=1)).max() >= 3, "the wild ensemble must exercise the fallback"


def test_two_million_columns_index_arithmetic(rcm, golden):
    """Maximum-size edge: 2,097,152 columns (32 x BASELINE's ensemble, 335 MB per state array, tau of 3.4 GB) on the
    10-wavelength table - the element offsets of tau pass 2^31, the byte offsets of the state arrays 2^28.  Replicated
    columns are bit-identical at the far end of every array, in the fused step, in the tau build and in the host-buffer
    step (chunk pipeline), and equal the 16-column run."""
    import gc
    base = 16
    rep = 131072
    ncol = base * rep
    s = rcm.Solver(0)
    s.set_repwvl_table_from(rcm.Table(table_path(10)))
    # (the split path's order of additions does not depend on the ensemble size: no tile shape to force)
    s.set_columns(golden["plevel"], golden["Tlayer"], golden["Tsurf"], golden["vmr9"], golden["rel_hum"])
    s.advance(2)
    small = s.get_state()
    tau_small, _, lt_small = s.build_tau()
    tile = lambda a: np.tile(a, (rep,) + (1,) * (a.ndim - 1))
    vmr9 = tile(golden["vmr9"])
    s.set_columns(golden["plevel"], tile(golden["Tlayer"]), tile(golden["Tsurf"]), vmr9, tile(golden["rel_hum"]))
    del vmr9
    gc.collect()
    sc = s.advance(2)
    big = s.get_state()
    for k in ("E_up", "E_down", "dE", "Tlayer", "Tsurf", "h2o", "dt", "time_h"):
        b = big[k].reshape(rep, base, -1)
        ref = small[k].reshape(base, -1)
        assert np.array_equal(b[0], ref) and np.array_equal(b[-1], ref) and np.array_equal(b[rep // 2 + 1], ref), k
    assert sc[-1, 2] == 0 and np.isfinite(sc).all()
    np.testing.assert_allclose(sc[-1, 0], rep * float(np.sum(float(golden["solar_irr"]) - small["E_up"][:, 0])), rtol=1e-12)
    del big
    gc.collect()
    tau, _, lt = s.build_tau()                       # [ncol][10][20] doubles = 3.4 GB: offsets beyond 2^31 elements
    t = tau.reshape(rep, base, 10, 20)
    assert np.array_equal(t[0], tau_small) and np.array_equal(t[-1], tau_small) and np.array_equal(t[rep - 7], tau_small)
    assert np.array_equal(lt.reshape(rep, base, 20)[-1], lt_small)
    del tau, t, lt
    gc.collect()
    s.close()


def test_full_size_properties(solver, rcm, golden):
    """BASELINE-size ensemble (65,536 columns x 100 wavelengths): size-independent properties.
    (1) replicated columns give bit-identical results wherever they sit in the ensemble;
    (2) energy bookkeeping: sum_l dE = solar + E_down[0] - E_up[0] to rounding;
    (3) member 0 (unperturbed) reproduces the reference's golden single-column fluxes."""
    ncol = 65536
    atm = rcm.read_atm(table_path(100).replace("Reduced100Forcing.rcmtab", "column21.atm"))
    pl = atm[:, 1]
    Tlev, vlev = rcm.make_ensemble(4096, 12345, pl, atm[:, 2], atm[:, 4:9].T.copy())
    st0 = rcm.init_columns(pl, Tlev, vlev)
    rep = ncol // 4096
    tile = lambda a: np.tile(a, (rep,) + (1,) * (a.ndim - 1))
    solver.set_repwvl_table_from(rcm.Table(table_path(100)))
    solver.set_columns(pl, tile(st0["Tlayer"]), np.full(ncol, 288.2), tile(st0["vmr9"]), tile(st0["rel_hum"]))
    solver.advance(1)
    st = solver.get_state()
    for k in ("E_up", "E_down", "dE", "Tlayer"):
        blocks = st[k].reshape(rep, 4096, -1)
        assert np.array_equal(blocks[0], blocks[-1]) and np.array_equal(blocks[0], blocks[rep // 2]), k
    lhs = st["dE"].sum(axis=1)
    rhs = golden["solar_irr"] + st["E_down"][:, 0] - st["E_up"][:, 0]
    assert np.max(np.abs(lhs - rhs)) < 1e-9
    assert relerr(st["E_up"][:1], golden["s1_E_up_100"][:1]) < RTOL
    assert np.all(np.isfinite(st["Tlayer"]))


def test_step_host_chunk_pipeline_is_bit_identical_to_the_resident_path(solver, rcm, golden):
    """rcm_step_host (host buffers, chunked copy/compute pipeline) == set_columns + advance + get_state, bit for bit,
    over two consecutive steps (the second one exercises the i > 0 branch: feedback + tau rebuild)."""
    atm = rcm.read_atm(table_path(100).replace("Reduced100Forcing.rcmtab", "column21.atm"))
    pl = atm[:, 1]
    ncol = 20000                                  # not a multiple of the chunk or tile size
    Tlev, vlev = rcm.make_ensemble(ncol, 77, pl, atm[:, 2], atm[:, 4:9].T.copy())
    st0 = rcm.init_columns(pl, Tlev, vlev)
    Ts0 = np.full(ncol, 288.2)
    solver.set_repwvl_table_from(rcm.Table(table_path(100)))
    solver.set_columns(pl, st0["Tlayer"], Ts0, st0["vmr9"], st0["rel_hum"])
    solver.advance(1)
    a1 = solver.get_state()
    solver.advance(1)
    a2 = solver.get_state()
    active = [k for k in range(9) if solver.params.species_mask >> k & 1]
    solver.set_columns(pl, st0["Tlayer"], Ts0, st0["vmr9"], st0["rel_hum"])
    b1 = solver.step_host(st0["Tlayer"], Ts0, st0["vmr9"][:, active, :])
    vmr1 = st0["vmr9"][:, active, :].copy()       # H2O after step 1 is unchanged (feedback acts from the 2nd iteration)
    b2 = solver.step_host(b1["Tlayer"], b1["Tsurf"], vmr1)
    for k in ("E_down", "E_up", "dE", "Tlayer", "Tsurf"):
        assert np.array_equal(a1[k], b1[k]), k
        assert np.array_equal(a2[k], b2[k]), k


def test_cplkavg_device_matches_reference(solver, golden_misc):
    m = golden_misc
    out = solver.cplkavg_device(m["cpl_lo"], m["cpl_hi"], m["cpl_T"])
    np.testing.assert_allclose(out, m["cpl_val"], rtol=1e-12)
    # the LBL kernel's variant (solver exp / division in the narrow-band Simpson branch, generic code otherwise)
    narrow = (1e7 / m["cpl_lo"] - 1e7 / m["cpl_hi"]) / (1e7 / m["cpl_lo"]) < 1e-2
    assert narrow.any() and (~narrow).any()
    solver.set_option(2, 1)
    out2 = solver.cplkavg_device(m["cpl_lo"], m["cpl_hi"], m["cpl_T"])
    solver.set_option(2, 0)
    np.testing.assert_allclose(out2, m["cpl_val"], rtol=1e-12)
    assert np.array_equal(out2[~narrow], out[~narrow])


def test_kernels_really_ran(solver):
    assert solver.launch_count() > 0
    ms, n = solver.kernel_time_ms()
    assert n > 0 and ms > 0
