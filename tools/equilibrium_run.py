"""BASELINE configs[3]: a 65,536-column synthetic ensemble, repwvl-100, stepped to equilibrium on one GPU.
Reports wall time, steps, the stationarity diagnostic and the ensemble-mean TOA imbalance (the reference's own
equilibrium has ~ +18 W/m2)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import our_first_climate_model_b200 as rcm
import bench

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
max_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-3
st = bench.build_ensemble(rcm, ncol, 12345)
p = rcm.default_params()
p.dT_converged = thr
s = rcm.Solver(0, p)
s.set_repwvl_table_from(rcm.Table(os.path.join(bench.GOLDEN, "Reduced100Forcing.rcmtab")))
s.set_columns(st["plevel"], st["Tlayer"], st["Tsurf"], st["vmr9"], st["rel_hum"])
t0 = time.perf_counter()
done, last = s.run_to_equilibrium(max_steps, 250)
s.synchronize()
wall = time.perf_counter() - t0
out = s.get_state()
print(f"ncol={ncol}: {done} steps in {wall:.2f} s ({1e3*wall/done:.3f} ms/step, {ncol*100*20*done/wall/1e9:.2f} G updates/s); "
      f"converged {int(last[2])}/{ncol} at {thr} K/step, max dT {last[1]:.2e} K, mean TOA net {last[0]/ncol:.3f} W/m2, "
      f"Tsurf mean {out['Tsurf'].mean():.3f} K (min {out['Tsurf'].min():.2f}, max {out['Tsurf'].max():.2f}), member 0 Tsurf {out['Tsurf'][0]:.4f} K")
