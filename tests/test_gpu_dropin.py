"""The literal drop-in: the reference driver main.cpp, unmodified, relinked against this library
(oracle/Makefile target _ref/main_b200).  Built in the container that has
/root/reference; the binary travels to the GPU box."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.parametrize("exe", ["main_b200_tau", "main_b200"])
def test_relinked_reference_driver_runs_on_the_gpu(exe, golden, tmp_path):
    path = os.path.join(BIN, exe)
    if not os.path.exists(path):
        pytest.skip("relinked driver not built (needs /root/reference at build time)")
    os.makedirs(tmp_path / "repwvl_V2.01_cpp")
    shutil.copy(os.path.join(GOLDEN, "column21.atm"), tmp_path / "repwvl_V2.01_cpp" / "test.atm")
    env = dict(os.environ, RCM_TABLE_DIR=GOLDEN, RCM_ADAPTER_LOG="1")
    r = subprocess.run([path], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = (tmp_path / "output.txt").read_text()
    # the rows the reference prints at t=0 (output.txt:8-27 of the reference)
    assert "solar irradiance: 236.882897" in out
    assert "0,25.000000,221.393000,635.177801,0.000000" in out and "19,975.000000,286.801000,288.883142,0.000000" in out
    taus = re.findall(r"read_tau: nwvl=(\d+) tau\[0\]\[19\]=(\S+)", r.stderr)
    assert len(taus) == 2 and all(int(n) == 100 for n, _ in taus)      # main.cpp:500 and :564 (n_steps = 1)
    assert float(taus[0][1]) == golden["tau100"][0, 0, 19]            # bit-identical optical depth
    if exe == "main_b200":
        olr = [float(x) for x in re.findall(r"radiative_transfer: OLR=(\S+)", r.stderr)]
        assert len(olr) == 2                                            # iterations i = 0 and 1
        assert abs(olr[0] - golden["s1_E_up_100"][0, 0]) < 1e-10 * olr[0]
        assert abs(olr[1] - golden["s5_trace_100"][0, 1, 22]) < 1e-9 * olr[1]


@pytest.mark.parametrize("prop_at_lev", [1, 0])
@pytest.mark.parametrize("n", [20, 100])
def test_read_tau_adapter_against_the_reference_build(prop_at_lev, n, golden, tmp_path):
    """One caller of read_tau (oracle/read_tau_cli.cpp, includes the reference's repwvl_thermal.h) linked against the
    reference's repwvl_thermal.cpp and against librcm_b200.so: bit-identical tau, wvl and weight - also for prop_at_Lev != 0
    (repwvl_thermal.cpp:219-224: T and VMRs given at the 21 levels), which the reference driver never uses."""
    ref, mine = (os.path.join(BIN, f"read_tau_cli_{k}") for k in ("ref", "b200"))
    if not (os.path.exists(ref) and os.path.exists(mine)):
        pytest.skip("read_tau_cli not built (needs /root/reference at build time)")
    rng = np.random.default_rng(7 + n)
    c = 3  # a perturbed member
    T = golden["Tlevel"][c] + rng.uniform(-4, 4, 21)
    vmr = np.zeros((9, 21))
    for k, row in zip((0, 2, 1, 5, 3), golden["vmr_ppm_level"][c]):  # file order H2O, O3, CO2, CH4, N2O -> read_tau's argument order
        vmr[k] = row * 1e-6
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(np.ascontiguousarray(golden["plevel"], dtype="<f8").tobytes() + T.astype("<f8").tobytes() + vmr.astype("<f8").tobytes())
    table = os.path.join(GOLDEN, f"Reduced{n}Forcing.rcmtab")
    out = {}
    for k, exe in (("ref", ref), ("mine", mine)):
        r = subprocess.run([exe, table, str(tmp_path / "in.bin"), str(tmp_path / f"{k}.bin"), str(prop_at_lev)],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (k, r.returncode, r.stderr)
        out[k] = (tmp_path / f"{k}.bin").read_bytes()
    nw = int(np.frombuffer(out["ref"][:4], dtype="<i4")[0])
    assert nw == n and len(out["ref"]) == 4 + 8 * (nw * 20 + 2 * nw)
    assert out["mine"] == out["ref"]
