"""The C++ ensemble RCE driver (csrc/host/rce_driver.cpp -> lib/rcm_rce): host code in the reference's language over the
C ABI.  Its output rows (the reference's output_conv format) must equal what the same run gives through the Python
binding; a run split by checkpoint / resume must give the same file."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, table_path

pytestmark = pytest.mark.gpu


def exe(rcm):
    p = os.path.join(os.path.dirname(rcm.library_path()), "rcm_rce")
    if not os.path.exists(p):
        rcm.build_library()
    return p


def run(rcm, *args):
    r = subprocess.run([exe(rcm), "--atm", os.path.join(GOLDEN, "column21.atm"), *args], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stdout


def test_driver_rows_equal_the_python_path_and_resume(rcm, tmp_path):
    ncol, seed, tab = 5, 12345, table_path(20)
    out = str(tmp_path / "output.txt")
    msg = run(rcm, "--table", tab, "--ncol", str(ncol), "--seed", str(seed), "--max-steps", "3", "--steps-exact", "--out", out)
    assert "5 columns, 3 iterations" in msg
    # the same run through the binding
    atm = rcm.read_atm(os.path.join(GOLDEN, "column21.atm"))
    pl = atm[:, 1].copy()
    Tlev, vlev = rcm.make_ensemble(ncol, seed, pl, atm[:, 2].copy(), atm[:, 4:9].T.copy())
    st = rcm.init_columns(pl, Tlev, vlev)
    p = rcm.default_params()
    p.dT_converged = 1e-3
    s = rcm.Solver(0, p)
    s.set_repwvl_table_from(rcm.Table(tab))
    s.set_columns(pl, st["Tlayer"], 288.2, st["vmr9"], st["rel_hum"])
    s.advance(3)
    got = s.get_state()
    s.close()
    ref = str(tmp_path / "ref.txt")
    rcm.write_profiles(ref, pl, got["Tlayer"], got["time_h"], header=True, column_ids=True)
    assert open(out).read() == open(ref).read()
    # two iterations, checkpoint, one more after a resume: the same file
    ck, out2 = str(tmp_path / "run.ckpt"), str(tmp_path / "output2.txt")
    run(rcm, "--table", tab, "--ncol", str(ncol), "--seed", str(seed), "--max-steps", "2", "--steps-exact", "--out", out2,
        "--checkpoint", ck)
    assert open(out2).read() != open(out).read()
    run(rcm, "--table", tab, "--resume", ck, "--max-steps", "1", "--steps-exact", "--out", out2)
    assert open(out2).read() == open(out).read()


def test_driver_reaches_the_reference_equilibrium(rcm, golden, tmp_path):
    """One column (the .atm file itself), Reduced100, 6,000 iterations: the surface temperature of the UNMODIFIED reference
    after the same 6,000 iterations (golden vector, member 0 = the file's column) within the 1e-3 K of north_star; rows in
    the reference's exact format (no column ids).  Then the stationarity stop."""
    out = str(tmp_path / "output.txt")
    msg = run(rcm, "--table", table_path(100), "--ncol", "1", "--max-steps", "6000", "--steps-exact", "--out", out)
    ts = float(msg.split("member 0: T_surface ")[1].split(" K")[0])
    assert abs(ts - float(golden["s6000_Tsurf_100"][0])) < 1e-3, msg
    lines = open(out).read().splitlines()
    assert lines[0] == "layer,player,Tlayer,theta,time" and len(lines) == 21 and lines[1].startswith("0,25.000000,")
    msg = run(rcm, "--table", table_path(100), "--ncol", "3", "--max-steps", "6000", "--check-every", "500", "--dT", "1e-3",
              "--out", out)
    assert "3/3 stationary" in msg and int(msg.split(" columns, ")[1].split(" iterations")[0]) < 6000


def test_driver_line_by_line_mode(rcm, tmp_path):
    """--lbl DIR: the five species tables written in the reference's text format (lbl.arts/README:5-11), read back by the
    drop-in of ASCII_file2xy2D inside the C++ driver, two iterations with 2xCO2 - same rows as the Python path."""
    lbl_atm = os.path.join(GOLDEN, "column21.lbl.atm")      # six columns: z p T air H2O O3 (lbl.arts/README:1-3)
    atm = rcm.read_atm(lbl_atm)
    assert atm.shape[1] == 6
    pl = atm[:, 1].copy()
    vbase = np.stack([atm[:, 4], atm[:, 5], np.full(21, 400.0), np.full(21, 1.7), np.full(21, 0.315)])  # README:13-16
    ncol, seed, nwvl = 6, 77, 700
    Tlev, vlev = rcm.make_ensemble(ncol, seed, pl, atm[:, 2].copy(), vbase)
    st = rcm.init_columns(pl, Tlev, vlev)
    h2o_ref, o3_ref = st["vmr9"][0, 0].copy(), st["vmr9"][0, 2].copy()
    wvl, tau5 = rcm.make_lbl_tables(nwvl, 777, pl, h2o_ref, o3_ref)
    d = tmp_path / "lbl"
    d.mkdir()
    for k, nm in enumerate(["h2o", "co2", "o3", "ch4", "n2o"]):
        rcm.write_lbl_asc(str(d / f"lbl.{nm}.asc"), wvl, tau5[k])
    out = str(tmp_path / "output.txt")
    r = subprocess.run([exe(rcm), "--atm", lbl_atm, "--lbl", str(d), "--co2-factor", "2", "--ncol", str(ncol), "--seed", str(seed),
                        "--max-steps", "2", "--steps-exact", "--out", out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    s = rcm.Solver(0)
    s.set_lbl_tables(wvl, tau5, h2o_ref, o3_ref, 2.0)
    s.set_columns(pl, st["Tlayer"], 288.2, st["vmr9"], st["rel_hum"])
    s.advance(2)
    got = s.get_state()
    s.close()
    ref = str(tmp_path / "ref.txt")
    rcm.write_profiles(ref, pl, got["Tlayer"], got["time_h"], header=True, column_ids=True)
    assert open(out).read() == open(ref).read()
    # a table with another wavelength grid is refused
    rcm.write_lbl_asc(str(d / "lbl.o3.asc"), wvl[:-1], tau5[2][:-1])
    r = subprocess.run([exe(rcm), "--atm", lbl_atm, "--lbl", str(d), "--max-steps", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "lbl.o3.asc" in r.stderr


def test_driver_argument_errors(rcm, tmp_path):
    r = subprocess.run([exe(rcm), "--atm", "/no/such.atm", "--table", table_path(10)], capture_output=True, text=True)
    assert r.returncode == 1 and "rcm_rce:" in r.stderr
    r = subprocess.run([exe(rcm), "--bogus"], capture_output=True, text=True)
    assert r.returncode == 1


def test_driver_on_two_gpus_writes_the_rows_of_the_one_gpu_run(rcm, tmp_path):
    """rcm_rce --gpus 2: one process, a thread and a solver per GPU, NCCL allreduce of the block scalars.  The profile rows
    are byte-identical to the 1-GPU run's (per-column results do not depend on the sharding), the run stops after the same
    number of iterations (the stationarity decision sees the same reduced scalars), and a checkpointed / resumed 2-GPU run
    ends in the same file."""
    if rcm.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    tab = table_path(100)
    common = ["--table", tab, "--ncol", "3001", "--seed", "7", "--check-every", "10", "--dT", "0.25"]
    o1, o2, o3 = (str(tmp_path / f"o{k}.txt") for k in (1, 2, 3))
    m1 = run(rcm, *common, "--max-steps", "200", "--out", o1)
    m2 = run(rcm, *common, "--max-steps", "200", "--out", o2, "--gpus", "2")
    it = lambda m: int(m.split(" iterations")[0].split()[-1])
    assert "3001 columns on 2 GPUs" in m2 and it(m1) == it(m2) and 20 <= it(m1) < 200, (m1, m2)
    assert open(o1).read() == open(o2).read()
    assert m1.split("iterations", 1)[1].split("mean TOA")[0] == m2.split("iterations", 1)[1].split("mean TOA")[0]
    ck = str(tmp_path / "two.ckpt")
    run(rcm, *common, "--max-steps", "10", "--steps-exact", "--out", o3, "--gpus", "2", "--checkpoint", ck)
    assert os.path.exists(ck + ".gpu0") and os.path.exists(ck + ".gpu1")
    run(rcm, *common, "--resume", ck, "--max-steps", str(it(m1) - 10), "--steps-exact", "--out", o3, "--gpus", "2")
    assert open(o3).read() == open(o1).read()
