"""The split path (rcm_split_kernels.cuh: (tile, wavelength split) units, fixed summation order) - the default of
rcm_advance - against the reference, against the fused tile kernel, and its defining property: results do not depend
on how the ensemble is cut into shards ("N-GPU result bitwise equal to 1-GPU result ... testable
with 1 GPU by running shards sequentially")."""
import numpy as np
import pytest

from conftest import table_path

pytestmark = pytest.mark.gpu
RTOL = 1e-10
KEYS = ("E_up", "E_down", "dE", "Tlayer", "Tsurf", "h2o", "dt", "time_h")
OPT_PATH = 5  # rcm_set_option: 0 = split path (default), 1 = fused tile kernel


def relerr(a, b):
    return float(np.max(np.abs(a - b) / np.max(np.abs(b), axis=-1, keepdims=True)))


def ensemble(rcm, ncol, seed, n=100):
    atm = rcm.read_atm(table_path(n).replace(f"Reduced{n}Forcing.rcmtab", "column21.atm"))
    pl = atm[:, 1].copy()
    Tlev, vlev = rcm.make_ensemble(ncol, seed, pl, atm[:, 2].copy(), atm[:, 4:9].T.copy())
    st = rcm.init_columns(pl, Tlev, vlev)
    st["plevel"], st["Tsurf"] = pl, Tlev[:, 20].copy()
    return st


def run(s, st, sl, steps):
    s.set_columns(st["plevel"], st["Tlayer"][sl], st["Tsurf"][sl], st["vmr9"][sl], st["rel_hum"][sl])
    sc = [s.advance(k) for k in steps]
    return s.get_state(), np.concatenate(sc)


@pytest.mark.parametrize("n", [10, 20, 100])
def test_split_path_against_the_reference_goldens(rcm, golden, n):
    s = rcm.Solver(0)
    s.set_repwvl_table_from(rcm.Table(table_path(n)))
    s.set_columns(golden["plevel"], golden["Tlayer"], golden["Tsurf"], golden["vmr9"], golden["rel_hum"])
    s.advance(1)
    st = s.get_state()
    scale = np.max(np.abs(golden[f"s1_E_up_{n}"]), axis=-1, keepdims=True)
    assert relerr(st["E_down"], golden[f"s1_E_down_{n}"]) < RTOL and relerr(st["E_up"], golden[f"s1_E_up_{n}"]) < RTOL
    assert float(np.max(np.abs(st["dE"] - golden[f"s1_dE_{n}"]) / scale)) < RTOL
    np.testing.assert_allclose(st["Tlayer"], golden[f"s1_Tlayer_{n}"], rtol=1e-11)
    s.advance(4)
    st = s.get_state()
    np.testing.assert_allclose(st["Tlayer"], golden[f"s5_Tlayer_{n}"], rtol=1e-10)
    np.testing.assert_allclose(st["h2o"], golden[f"s5_h2o_{n}"], rtol=1e-10)
    np.testing.assert_allclose(st["time_h"], golden[f"s5_time_h_{n}"], rtol=1e-6)
    s.close()


@pytest.mark.parametrize("n,ncol", [(100, 333), (20, 50), (10, 17)])
def test_split_path_equals_the_fused_kernel_to_rounding(rcm, n, ncol):
    """Same arithmetic per (column, wavelength, angle), another order of the spectral sum: <= 1e-13 relative over six
    steps - and not bit-identical (the two paths really are different code)."""
    st = ensemble(rcm, ncol, 5 + n, n)
    s = rcm.Solver(0)
    s.set_repwvl_table_from(rcm.Table(table_path(n)))
    out = {}
    for path in (0, 1):
        s.set_option(OPT_PATH, path)
        out[path], sc = run(s, st, slice(None), (1, 2, 3))
        assert np.isfinite(sc).all() and sc.shape == (6, 4)
    s.close()
    for k in ("E_up", "E_down", "Tlayer", "Tsurf", "h2o", "dt"):
        np.testing.assert_allclose(out[0][k], out[1][k], rtol=1e-12, atol=0, err_msg=k)
    assert relerr(out[0]["E_up"], out[1]["E_up"]) < 1e-13
    if n == 100:
        assert not np.array_equal(out[0]["E_up"], out[1]["E_up"])


@pytest.mark.parametrize("shards", [2, 8, 7])
def test_shards_are_bit_identical_to_the_whole_ensemble(rcm, shards):
    """One 20,000-column ensemble stepped as a whole and as 2 / 8 / 7 sequential shards (contiguous blocks as
    distributed.shard_range deals them; 20,000 / 7 and / 8 are not multiples of the 16-column tile): every per-column
    output is bit-identical - fluxes, heating rates, temperatures, time step - over three steps (initial-profile tau,
    then the feedback / re-sort branch), and so are max-type ensemble scalars."""
    from our_first_climate_model_b200.distributed import shard_range
    ncol = 20000
    st = ensemble(rcm, ncol, 31)
    s = rcm.Solver(0)
    s.set_repwvl_table_from(rcm.Table(table_path(100)))
    whole, sc_whole = run(s, st, slice(None), (1, 2))
    parts, scs = [], []
    for r in range(shards):
        lo, hi = shard_range(ncol, r, shards)
        o, sc = run(s, st, slice(lo, hi), (1, 2))
        parts.append(o)
        scs.append(sc)
    s.close()
    for k in KEYS:
        assert np.array_equal(np.concatenate([p[k] for p in parts]), whole[k]), k
    scs = np.array(scs)
    assert np.array_equal(scs[:, :, 1].max(axis=0), sc_whole[:, 1]) and np.array_equal(scs[:, :, 3].max(axis=0), sc_whole[:, 3])
    np.testing.assert_allclose(scs[:, :, 0].sum(axis=0), sc_whole[:, 0], rtol=1e-12)   # sums: other order over the ranks


def test_one_column_equals_the_same_column_inside_a_big_ensemble(rcm):
    """Extreme case of the above: a column alone (1 tile, 5 units) and at position 12,345 of 20,000."""
    st = ensemble(rcm, 20000, 31)
    s = rcm.Solver(0)
    s.set_repwvl_table_from(rcm.Table(table_path(100)))
    whole, _ = run(s, st, slice(None), (3,))
    one, _ = run(s, st, slice(12345, 12346), (3,))
    s.close()
    for k in KEYS:
        assert np.array_equal(one[k][0], whole[k][12345]), k


def test_step_host_pipeline_on_the_split_path(rcm):
    """rcm_step_host's chunk pipeline (three streams, own work counter each) == resident stepping, bit for bit."""
    ncol = 40000
    st = ensemble(rcm, ncol, 8)
    s = rcm.Solver(0)
    s.set_repwvl_table_from(rcm.Table(table_path(100)))
    s.set_columns(st["plevel"], st["Tlayer"], st["Tsurf"], st["vmr9"], st["rel_hum"])
    s.advance(1)
    a1 = s.get_state()
    s.advance(1)
    a2 = s.get_state()
    active = [k for k in range(9) if s.params.species_mask >> k & 1]
    s.set_columns(st["plevel"], st["Tlayer"], st["Tsurf"], st["vmr9"], st["rel_hum"])
    b1 = s.step_host(st["Tlayer"], st["Tsurf"], st["vmr9"][:, active, :])
    b2 = s.step_host(b1["Tlayer"], b1["Tsurf"], st["vmr9"][:, active, :])
    for k in ("E_down", "E_up", "dE", "Tlayer", "Tsurf"):
        assert np.array_equal(a1[k], b1[k]) and np.array_equal(a2[k], b2[k]), k
    s.close()


def test_perturbed_members_reach_the_reference_equilibrium(rcm, golden, golden_eq):
    """north_star: "the equilibrium temperature profile within 1e-3 K" - for PERTURBED ensemble members (13 of the 16
    columns; two more sit exactly on table nodes), 6,000 iterations of main.cpp:531-583 by the unmodified reference
    (tests/golden/ref_equilibrium.npz) against the GPU run in blocks of 250 fused steps."""
    s = rcm.Solver(0)
    s.set_repwvl_table_from(rcm.Table(table_path(100)))
    s.set_columns(golden["plevel"], golden["Tlayer"], golden["Tsurf"], golden["vmr9"], golden["rel_hum"])
    n = int(golden_eq["nsteps"])
    sc = None
    for _ in range(n // 250):
        sc = s.advance(250)
    st = s.get_state()
    s.close()
    dT = np.abs(st["Tlayer"] - golden_eq["Tlayer"]).max(axis=1)
    assert dT.max() < 1e-3, dT
    assert np.abs(st["Tsurf"] - golden_eq["Tsurf"]).max() < 1e-3
    # far tighter in practice: the trajectory is stable
    assert dT.max() < 1e-6 and relerr(st["E_up"], golden_eq["E_up"]) < 1e-9
    np.testing.assert_allclose(st["h2o"], golden_eq["h2o"], rtol=1e-8)
    assert sc[-1, 1] < 1e-2 and golden_eq["Tsurf"].max() - golden_eq["Tsurf"].min() > 10.0  # members really differ


def test_step_host_graph_path_is_bit_identical(rcm):
    """From the second call on rcm_step_host launches its whole chunk pipeline as ONE captured CUDA graph (page-locked host
    buffers; the first call and calls with pageable buffers stay direct).  Five consecutive calls - direct, capture +
    launch, three relaunches - against resident stepping, bit for bit; then other output buffers (re-capture)."""
    import torch
    ncol = 40000
    st = ensemble(rcm, ncol, 9)
    s = rcm.Solver(0)
    s.set_repwvl_table_from(rcm.Table(table_path(100)))
    s.set_columns(st["plevel"], st["Tlayer"], st["Tsurf"], st["vmr9"], st["rel_hum"])
    ref = []
    for _ in range(6):
        s.advance(1)
        ref.append(s.get_state())
    pin = lambda *shape: torch.empty(*shape, dtype=torch.float64).pin_memory()
    T_in, Ts_in, v_in = pin(ncol, 20), pin(ncol), pin(ncol, s.nactive, 20)
    outs = [[pin(ncol, 21), pin(ncol, 21), pin(ncol, 20), pin(ncol, 20), pin(ncol)] for _ in range(2)]
    active = [k for k in range(9) if s.params.species_mask >> k & 1]
    T_in.copy_(torch.from_numpy(st["Tlayer"]))
    Ts_in.copy_(torch.from_numpy(st["Tsurf"]))
    v_in.copy_(torch.from_numpy(np.ascontiguousarray(st["vmr9"][:, active, :])))
    s.set_columns(st["plevel"], st["Tlayer"], st["Tsurf"], st["vmr9"], st["rel_hum"])
    l0 = s.launch_count()
    for k in range(6):
        o = outs[0] if k < 5 else outs[1]                      # the sixth call writes to other buffers: re-capture
        ptrs = [T_in.data_ptr(), Ts_in.data_ptr(), v_in.data_ptr() if k == 0 else 0] + [t.data_ptr() for t in o]
        s.step_host_ptrs(*ptrs)
        for name, t in zip(("E_down", "E_up", "dE", "Tlayer", "Tsurf"), o):
            assert np.array_equal(t.numpy(), ref[k][name]), (k, name)
        T_in.copy_(o[3])                                          # the next call continues from this call's result
        Ts_in.copy_(o[4])
    assert s.launch_count() - l0 >= 6 * (3 * 8 + 1)               # graph launches count their kernels too
    assert s.host_graph_stats() == (2, 5)                         # captured at call 2 and call 6, replayed 5 times
    s.close()


@pytest.mark.parametrize("ncol,n", [(333, 100), (5000, 100), (50, 20), (17, 10)])
def test_multi_step_launch_is_bit_identical_to_three_launches_per_step(rcm, ncol, n):
    """rcm_advance(k >= 2) as ONE persistent launch (per-tile step flags, the K5 body run by the CTA that completes a tile's
    step; option 6 = 2 forces it) against a K5 / unit / K5 launch sequence per step (option 6 = 0): every output and every
    step's ensemble scalars, over blocks of different lengths - and against single steps."""
    st = ensemble(rcm, ncol, 77 + n, n)
    s = rcm.Solver(0)
    s.set_repwvl_table_from(rcm.Table(table_path(n)))
    out = {}
    for multi in (2, 0):
        s.set_option(6, multi)
        s.set_columns(st["plevel"], st["Tlayer"], st["Tsurf"], st["vmr9"], st["rel_hum"])
        l0 = s.launch_count()
        sc = [s.advance(k) for k in (2, 7, 1, 40)]
        launches = s.launch_count() - l0
        out[multi] = (s.get_state(), np.concatenate(sc))
        # one launch per block: K5 prep + unit kernel + scalar reduction per call (the single step: prep, unit, finish, reduction)
        # (+ the two table kernels after the first rcm_set_columns)
        assert (launches <= 3 * 3 + 4 + 2) if multi else (launches >= 4 * 1 + 2 * 50 + 4), (multi, launches)
    s.set_option(6, 0)
    single = run(s, st, slice(None), (1,) * 50)
    s.close()
    for k in KEYS:
        assert np.array_equal(out[2][0][k], out[0][0][k]), k
        assert np.array_equal(out[2][0][k], single[0][k]), k
    assert np.array_equal(out[2][1], out[0][1]) and np.array_equal(out[2][1], single[1])
    assert np.isfinite(out[2][1]).all() and out[2][1].shape == (50, 4)
