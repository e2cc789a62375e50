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
