"""Step time of the repwvl kernel with 1, 2 and 3 resident CTAs per SM (RCM_CTAS_PER_SM limits the grid): how much of the
FP64 pipe one, two and three warps per SM sub-partition can keep busy.  Usage: python tools/occupancy_probe.py [ncol]"""
import os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
ncol = sys.argv[1] if len(sys.argv) > 1 else "65536"
code = r'''
import os, sys, time
sys.path.insert(0, %r)
import numpy as np, bench, our_first_climate_model_b200 as rcm
ncol = int(%r)
st = bench.build_ensemble(rcm, ncol, 12345)
s = rcm.Solver(0)
s.set_repwvl_table_from(rcm.Table(os.path.join(bench.GOLDEN, "Reduced100Forcing.rcmtab")))
s.set_columns(st["plevel"], st["Tlayer"], st["Tsurf"], st["vmr9"], st["rel_hum"])
s.advance(3, want_scalars=False); s.synchronize(); s.kernel_time_ms(reset=True)
for _ in range(10): s.advance_async(1)
s.synchronize()
ms, n = s.kernel_time_ms(reset=True)
print("RCM_CTAS_PER_SM=%%s: %%.3f ms per step (%%d launches), %%.2f G updates/s" %% (os.environ.get("RCM_CTAS_PER_SM", "3"), ms, n, ncol * 2000 / ms / 1e6))
''' % (ROOT, ncol)
for k in ("1", "2", "3"):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, RCM_CTAS_PER_SM=k))
