"""Host side of the drop-in boundary (no GPU): the C-ABI library, loaders, initialisation."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, table_path
from oracle import refcpu as R

REF = "/root/reference"
needs_refsrc = pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not mounted")
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")


def test_library_loads_and_exports_every_declared_symbol(rcm):
    lib = rcm.load_library()
    assert len(rcm.DECLARED_SYMBOLS) >= 35
    missing = [s for s in rcm.DECLARED_SYMBOLS if not hasattr(lib, s)]
    assert not missing, missing
    # and nothing torch-typed in the boundary: the header is plain C
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "rcm_b200.h")])


def test_no_cpu_fallback(rcm):
    if rcm.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(rcm.RcmError, match="no CUDA device"):
        rcm.Solver(0)


def test_product_never_imports_the_oracle():
    """The shipped path must not route through the CPU oracle (or any CPU fallback)."""
    import re
    pkg = os.path.join(ROOT, "our_first_climate_model_b200")
    pat = re.compile(r"(from|import)\s+oracle|oracle[/\\.](port|refcpu|rcm_oracle|_ref)|liboracle|libref_oracle|rcmo_")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not pat.search(txt), f"{f} references the oracle"


def test_table_loader_rcmtab(rcm, port):
    for n in (10, 20, 100):
        t = rcm.Table(table_path(n))
        ref = port.load_rcmtab(table_path(n))
        assert (t.n_tpert, t.n_species, t.n_wvl, t.n_p) == (9, 9, n, 41)
        for nm in ("xsec", "wvl", "weight", "p_grid", "t_ref", "t_pert", "vmrs_ref"):
            assert np.array_equal(getattr(t, nm), ref[nm]), nm
    t = rcm.Table(table_path(100))
    assert t.p_grid[0] == 110000.0 and t.p_grid[0] > t.p_grid[-1]          # descending
    assert list(t.t_pert) == [-120, -70, -40, -20, 0, 20, 40, 70, 120]
    assert t.weight.min() < 0 < t.weight.max()                              # signed weights


@needs_refsrc
def test_table_loader_netcdf4_subset(rcm):
    """The C++ HDF5-subset reader on the reference's own .nc files == the independent Python reader."""
    for n in (10, 20, 100):
        a = rcm.Table(f"{REF}/repwvl_V2.01_cpp/Reduced{n}Forcing.nc")
        b = rcm.Table(table_path(n))
        for nm in ("xsec", "wvl", "weight", "p_grid", "t_ref", "t_pert", "vmrs_ref"):
            assert np.array_equal(getattr(a, nm), getattr(b, nm)), nm


def test_table_loader_errors(rcm, tmp_path):
    with pytest.raises(rcm.RcmError, match="not found"):
        rcm.Table(str(tmp_path / "missing.nc"))
    bad = tmp_path / "bad.nc"
    bad.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 200)
    with pytest.raises(rcm.RcmError, match="format"):
        rcm.Table(str(bad))
    trunc = tmp_path / "trunc.rcmtab"
    trunc.write_bytes(open(table_path(10), "rb").read()[:5000])
    with pytest.raises(rcm.RcmError, match="format"):
        rcm.Table(str(trunc))


def test_atm_reader(rcm):
    a = rcm.read_atm(os.path.join(GOLDEN, "column21.atm"))
    assert a.shape == (21, 9)
    assert a[0, 1] == 0 and a[-1, 1] == 1000 and a[-1, 2] == 288.2 and a[0, 4] == 5.246 and a[5, 8] == 0.315
    b = rcm.read_atm(os.path.join(GOLDEN, "column21.lbl.atm"))
    assert b.shape == (21, 6) and np.array_equal(a[:, :6], b)
    if os.path.isdir(REF):
        assert np.array_equal(rcm.read_atm(f"{REF}/repwvl_V2.01_cpp/test.atm"), a)


def test_init_columns_and_solar_match_oracle_bitwise(rcm, port, golden):
    g = golden
    a = rcm.init_columns(g["plevel"], g["Tlevel"], g["vmr_ppm_level"], 1.0)
    b = port.init_columns(g["plevel"], g["Tlevel"], g["vmr_ppm_level"], 1.0)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["Tlayer"], g["init_Tlayer"]) and np.array_equal(a["rel_hum"], g["init_rel_hum"])
    a2 = rcm.init_columns(g["plevel"], g["Tlevel"][:2], g["vmr_ppm_level"][:2], 2.0)   # 2xCO2 (main.cpp:452-456)
    assert np.array_equal(a2["vmr9"][:, 1], 2.0 * 1e-6 * 400 * np.ones((2, 20)))
    assert rcm.solar_setup()["solar_irr"] == float(g["solar_irr"])
    p = rcm.default_params()
    assert (p.nangle, p.cloud_layer, p.cloud_tau, p.dp, p.max_dT, p.dt_cap, p.species_mask) == (30, 17, 1.0, 50.0, 5.0,
                                                                                               43200.0, 0x2F)


def test_lowerpos_edge_semantics(rcm, golden_misc, port):
    asc, desc = [0.0, 1.0, 2.0, 3.0], [3.0, 2.0, 1.0, 0.0]
    assert [rcm.lowerpos(asc, x) for x in golden_misc["lp_x"]] == list(golden_misc["lp_asc"])
    assert [rcm.lowerpos(desc, x) for x in golden_misc["lp_x"]] == list(golden_misc["lp_desc"])
    rng = np.random.default_rng(0)
    t = rcm.Table(table_path(100))
    for x in rng.uniform(-100, 2e5, 300):
        assert rcm.lowerpos(t.p_grid, x) == port.lowerpos(t.p_grid, x)


def test_ensemble_is_deterministic_and_member0_is_the_base_column(rcm, golden):
    pl, bT, bv = golden["plevel"], golden["Tlevel"][0], golden["vmr_ppm_level"][0]
    T1, v1 = rcm.make_ensemble(14, 12345, pl, bT, bv)
    assert np.array_equal(T1, golden["Tlevel"]) and np.array_equal(v1, golden["vmr_ppm_level"])
    assert np.array_equal(T1[0], bT) and np.array_equal(v1[0], bv)
    T2, _ = rcm.make_ensemble(200, 7, pl, bT, bv)
    assert np.abs(T2 - bT).max() < 20 and np.abs(T2[1:] - bT).min(axis=1).max() > 0
    Ta, _ = rcm.make_ensemble(50, 7, pl, bT, bv)
    assert np.array_equal(Ta, T2[:50])  # a shard of a bigger ensemble is the same columns


def _write(path, text):
    with open(path, "w") as f:
        f.write(text)
    return str(path)


def test_ascii_reader_semantics(rcm, tmp_path):
    p = _write(tmp_path / "a.asc", "# comment\n% another\n\n 4000.5 1e-3 2e-3   3e-3\n\t5000 0.1 0.2 0.3 # trailing\n"
                                   "6000 x 0.5 1.5e2\n")
    st, x, y = rcm.ascii_file2xy2D(p)
    assert st == 0 and list(x) == [4000.5, 5000.0, 6000.0]
    assert np.array_equal(y, [[1e-3, 2e-3, 3e-3], [0.1, 0.2, 0.3], [0.0, 0.5, 150.0]])   # non-numeric -> 0
    assert rcm.ascii_file2xy2D(str(tmp_path / "nope.asc"))[0] == -1                       # ASCIIFILE_NOT_FOUND
    assert rcm.ascii_file2xy2D(_write(tmp_path / "r.asc", "1 2 3\n4 5\n"))[0] == -5       # NOT_A_RECTANGULAR_MATRIX
    assert rcm.ascii_file2xy2D(_write(tmp_path / "e.asc", "# only comments\n\n"))[0] == -5  # empty: min != max


@needs_ref
def test_ascii_reader_equals_reference_reader(rcm, tmp_path):
    rng = np.random.default_rng(3)
    nw = 500
    wvl = np.sort(rng.uniform(4000, 1e5, nw))
    tau = 10 ** rng.uniform(-8, 3, (nw, 20))
    lines = ["# wavelength[nm] dtau(20 layers, top-down)  (lbl.arts/README format)"]
    for i in range(nw):
        sep = "\t" if i % 3 == 0 else " "
        lines.append(sep.join([f"{wvl[i]:.6f}"] + [f"{v:.9e}" for v in tau[i]]) + ("   " if i % 5 == 0 else ""))
        if i % 97 == 0:
            lines.append("")
            lines.append("% block comment")
    p = _write(tmp_path / "lbl.co2.asc", "\n".join(lines) + "\n")
    st1, x1, y1 = R.ascii_file2xy2D(p)
    st2, x2, y2 = rcm.ascii_file2xy2D(p)
    assert st1 == st2 == 0 and np.array_equal(x1, x2) and np.array_equal(y1, y2) and y2.shape == (nw, 20)
    for bad in ("1 2 3\n4 5\n", "", "#x\n"):
        pb = _write(tmp_path / "bad.asc", bad)
        assert R.ascii_file2xy2D(pb)[0] == rcm.ascii_file2xy2D(pb)[0]


def test_cplkavg_host_matches_reference_samples(rcm, golden_misc):
    m = golden_misc
    got = np.array([rcm.cplkavg_host(a, b, t)[0] for a, b, t in zip(m["cpl_lo"], m["cpl_hi"], m["cpl_T"])])
    assert np.array_equal(got, m["cpl_val"])
    assert rcm.cplkavg_host(500.0, 400.0, 300.0)[1] == 1


def test_write_profiles_reproduces_reference_output_rows(rcm, golden, tmp_path):
    """rcm_write_profiles = output_conv (main.cpp:102-114) batched: the t=0 rows of the reference's committed output.txt
    (/root/reference/output.txt:7-27) byte for byte; append / header / column-id forms."""
    rows = """0,25.000000,221.393000,635.177801 1,75.000000,221.393000,464.060873 2,125.000000,221.393000,401.041758
    3,175.000000,221.393000,364.282852 4,225.000000,221.393000,339.042853 5,275.000000,225.299500,325.799855
    6,325.000000,232.616500,320.702550 7,375.000000,239.063000,316.386335 8,425.000000,244.842000,312.651495
    9,475.000000,250.091000,309.365099 10,525.000000,254.908000,306.434706 11,575.000000,259.365500,303.793542
    12,625.000000,263.518000,301.391001 13,675.000000,267.409000,299.189516 14,725.000000,271.073000,297.159548
    15,775.000000,274.537000,295.276557 16,825.000000,277.824000,293.521596 17,875.000000,280.953500,291.879487
    18,925.000000,283.941500,290.337186 19,975.000000,286.801000,288.883142""".split()
    st = rcm.init_columns(golden["plevel"], golden["Tlevel"][:1], golden["vmr_ppm_level"][:1])
    path = str(tmp_path / "output.txt")
    rcm.write_profiles(path, golden["plevel"], st["Tlayer"], 0.0, header=True)
    want = "layer,player,Tlayer,theta,time\n" + "".join(r + ",0.000000\n" for r in rows)
    assert open(path).read() == want
    rcm.write_profiles(path, golden["plevel"], st["Tlayer"], np.float32(3.668250), append=True)   # dt/3600 of step 0
    txt = open(path).read()
    assert txt.startswith(want) and txt.endswith("19,975.000000,286.801000,288.883142,3.668250\n") and txt.count("\n") == 41
    # an ensemble: 3000 columns x 20 rows with column ids, large enough to cross the writer's flush threshold
    n = 3000
    T = np.tile(st["Tlayer"], (n, 1)) + np.arange(n)[:, None] * 1e-3
    rcm.write_profiles(path, golden["plevel"], T, np.arange(n, dtype=np.float32), header=True, column_ids=True)
    lines = open(path).read().splitlines()
    assert lines[0] == "column,layer,player,Tlayer,theta,time" and len(lines) == 1 + 20 * n
    c, l = 2999, 7
    player = (golden["plevel"][l] + golden["plevel"][l + 1]) / 2
    conv = (1000.0 / player) ** (2.0 / 7.0)
    assert lines[1 + 20 * c + l] == "%d,%d,%f,%f,%f,%f" % (c, l, player, T[c, l], T[c, l] * conv, float(c))
    with pytest.raises(rcm.RcmError):
        rcm.write_profiles(str(tmp_path / "no_such_dir" / "x.txt"), golden["plevel"], st["Tlayer"], 0.0)


# ---- exact-signature boundary of the line-by-line side (include/rcm_b200_adapters.hpp) ------------------
BIN = os.path.join(ROOT, "oracle", "_ref")


def _needs_bin(name):
    path = os.path.join(BIN, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not built (needs /root/reference at build time)")
    return path


def test_unmodified_testlblarts_links_against_the_product_library(rcm, tmp_path):
    """The reference's only caller of ASCII_file2xy2D (lbl.arts/testlblarts.cpp:13-38), compiled UNMODIFIED and linked
    against librcm_b200.so instead of lbl.arts/ascii.cpp, reads a table in the README's format."""
    exe = _needs_bin("testlblarts_b200")
    nw = 321
    os.makedirs(tmp_path / "lbl.arts")
    wvl = np.linspace(4000.0, 1e5, nw)
    tau = np.random.default_rng(0).uniform(0, 3, (nw, 20))
    rcm.write_lbl_asc(str(tmp_path / "lbl.arts" / "lbl.co2.asc"), wvl, tau)
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert f" ... read {nw} wavelengths and 20 layers from ./lbl.arts/lbl.co2.asc" in r.stderr   # testlblarts.cpp:32-33
    os.remove(tmp_path / "lbl.arts" / "lbl.co2.asc")
    r = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert r.returncode == 255 and "Error -1 reading ./lbl.arts/lbl.co2.asc" in r.stderr          # ASCIIFILE_NOT_FOUND


def test_ascii_file2xy2D_exact_signature_and_ownership(rcm, tmp_path):
    """extern "C" int ASCII_file2xy2D(char*, int*, int*, double**, double***) (ascii.h:63): x is one calloc'ed vector, y an
    array of nx calloc'ed rows released by ASCII_free_double(y, nx) (ascii.cpp:955-965, :1612-1613)."""
    lib = rcm.load_library()
    p = _write(tmp_path / "t.asc", "# c\n1 10 20 30\n2 11 21 31\n\n3 12 22 x\n")
    nx, ny = ctypes.c_int(0), ctypes.c_int(0)
    x = ctypes.POINTER(ctypes.c_double)()
    y = ctypes.POINTER(ctypes.POINTER(ctypes.c_double))()
    st = lib.ASCII_file2xy2D(ctypes.c_char_p(p.encode()), ctypes.byref(nx), ctypes.byref(ny), ctypes.byref(x), ctypes.byref(y))
    assert st == 0 and (nx.value, ny.value) == (3, 3)
    assert [x[i] for i in range(3)] == [1.0, 2.0, 3.0]
    assert [[y[i][j] for j in range(3)] for i in range(3)] == [[10, 20, 30], [11, 21, 31], [12, 22, 0]]
    assert lib.ASCII_free_double(y, nx) == 0
    ctypes.CDLL(None).free(x)
    for bad, code in (("1 2 3\n4 5\n", -5), ("", -5)):
        assert lib.ASCII_file2xy2D(_write(tmp_path / "b.asc", bad).encode(), ctypes.byref(nx), ctypes.byref(ny),
                                   ctypes.byref(x), ctypes.byref(y)) == code
    assert lib.ASCII_file2xy2D(str(tmp_path / "missing").encode(), ctypes.byref(nx), ctypes.byref(ny), ctypes.byref(x),
                               ctypes.byref(y)) == -1


def test_cplkavg_exact_signature_against_the_reference_build(golden_misc):
    """double cplkavg(double, double, double) (cplkavg.h:7): one caller (oracle/cplkavg_cli.cpp, includes the reference's
    header) linked against the reference's cplkavg.cpp and against librcm_b200.so gives the same numbers; bad arguments
    end the process with the reference's message and exit status (cplkavg.cpp:144-146, :32-35)."""
    ref, mine = _needs_bin("cplkavg_cli_ref"), _needs_bin("cplkavg_cli_b200")
    m = golden_misc
    inp = "".join(f"{a!r} {b!r} {t!r}\n" for a, b, t in zip(m["cpl_lo"].tolist(), m["cpl_hi"].tolist(), m["cpl_T"].tolist()))
    out = {}
    for k, exe in (("ref", ref), ("mine", mine)):
        r = subprocess.run([exe], input=inp, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        out[k] = np.array([float(v) for v in r.stdout.split()])
    assert out["ref"].shape == m["cpl_val"].shape and np.array_equal(out["ref"], m["cpl_val"])
    assert np.array_equal(out["mine"], out["ref"])
    for exe in (ref, mine):
        r = subprocess.run([exe], input="500 400 300\n", capture_output=True, text=True, timeout=60)
        assert r.returncode == 1 and "planck_func1--temperature or wavenums. wrong" in r.stderr


def test_roofline_constants_belong_to_the_committed_kernels():
    """bench.py computes `roofline.frac` from FP64-instruction counts taken out of ncu captures (profiles/roofline_capture.json).
    The captures are only valid for the kernel sources they were taken on: the json records their sha256, and this test fails
    when a kernel source has moved since - re-capture (tools/profile_r2.sh) and run tools/update_roofline_capture.py."""
    import bench
    cap = bench.load_capture()
    assert not cap["stale"], ("kernel sources changed after the ncu capture recorded in profiles/roofline_capture.json: "
                              f"recorded {cap.get('sources_sha256')}, now {cap['current_sha']}")
    for w in ("step", "lbl"):
        assert 200 < cap[w]["exec_fp64_per_unit"] < 600 and cap[w]["kernel"].find("rcm_") >= 0
    assert "rcm_split_rt_kernel" in cap["step"]["kernel"] and cap["step"]["ncol"] == 65536 and cap["step"]["nwvl"] == 100


def test_clock_sampler_evaluates_only_the_timed_window():
    """bench.ClockSampler: one nvidia-smi for all GPUs, started before warm-up; only rows that arrived inside the timed
    window count, per GPU, and throttle reasons of other GPUs / other times do not leak in."""
    import bench

    class FakeProc:
        def terminate(self):
            pass

    s = bench.ClockSampler(2)
    s.proc = FakeProc()
    row = lambda idx, sm, *flags: [str(idx), str(sm), "1965", "700.0"] + [("Active" if f else "Not Active") for f in flags]
    s.rows = [(10.00, row(0, 345, 0, 0, 0, 0)),                      # idle, before the window
              (10.02, row(0, 1965, 1, 0, 0, 0)),                     # hw_slowdown before the window: must not count
              (11.00, row(0, 1965, 0, 0, 0, 0)), (11.00, row(1, 1950, 0, 0, 0, 1)), (11.00, row(2, 300, 0, 1, 0, 0)),
              (11.05, row(0, 1965, 0, 0, 0, 0)), (11.05, row(1, 1935, 0, 0, 0, 1)),
              (12.00, row(1, 210, 0, 0, 1, 0))]                      # after the window
    s.window(10.99, 11.06)
    c = s.stop()
    assert c["samples"] == 4 and c["gpus_sampled"] == 2 and c["reasons"] == ["sw_power_cap"]
    assert c["sm_max_mhz"] == 1965.0 and c["per_gpu_sm_mhz"] == {"min": 1942.5, "max": 1965.0}


def test_multi_step_protocol_model():
    """A host-side model of rcm_split_multi_kernel's scheduling (rcm_split_unit_loop.inc, RCM_SPLIT_MULTI): items = (step, unit) in
    step-major order from one counter, a unit of step n may start when ready[tile] >= n, the CTA that completes a tile's last
    unit of a step runs the tile's finish / prep and publishes ready[tile] = n + 1.  Under random interleavings of any number
    of CTAs - fewer than items, more than items, one - every item completes (dependencies point to lower item numbers only),
    each (tile, step) is finished exactly once, after all of its units and before any unit of the next step."""
    import random
    rng = random.Random(5)
    for ntiles, nsplit, nsteps, nctas in [(7, 5, 6, 4), (3, 5, 4, 40), (16, 1, 5, 9), (5, 2, 9, 1), (32, 5, 3, 13)]:
        nunits = ntiles * nsplit
        total = nunits * nsteps
        counter, done, ready = 0, [0] * ntiles, [0] * ntiles
        log = []                                            # ("unit", tile, step) / ("k5", tile, step) in completion order
        state = [("take", None)] * nctas                    # per CTA: take an item / wait for its flag / work / exited
        exited = 0
        for _ in range(200 * total + 1000):
            if exited == nctas:
                break
            c = rng.randrange(nctas)
            kind, item = state[c]
            if kind == "take":
                item, counter = counter, counter + 1
                state[c] = ("exit", None) if item >= total else ("wait", item)
            elif kind == "wait":
                step, tile = item // nunits, (item % nunits) // nsplit
                if ready[tile] >= step:
                    state[c] = ("work", item)
            elif kind == "work":
                step, tile = item // nunits, (item % nunits) // nsplit
                log.append(("unit", tile, step))
                done[tile] += 1
                if done[tile] % nsplit == 0:               # this CTA finished the tile's step: K5 body in place, then publish
                    log.append(("k5", tile, step))
                    ready[tile] = step + 1
                state[c] = ("take", None)
            elif kind == "exit":
                state[c] = ("gone", None)
                exited += 1
        assert exited == nctas, "the model deadlocked"
        assert sum(1 for e in log if e[0] == "unit") == total
        pos = {e: i for i, e in enumerate(log)}
        for tile in range(ntiles):
            for step in range(nsteps):
                k5 = pos[("k5", tile, step)]
                units = [i for i, e in enumerate(log) if e == ("unit", tile, step)]
                assert len(units) == nsplit and max(units) < k5
                if step + 1 < nsteps:
                    assert k5 < min(i for i, e in enumerate(log) if e == ("unit", tile, step + 1))
