"""Timing of the line-by-line path (BASELINE configs 3 / 5 shape): ncol columns x nwvl synthetic LBL wavelengths."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import our_first_climate_model_b200 as rcm

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nwvl = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
co2 = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
atm = rcm.read_atm(os.path.join(G, "column21.lbl.atm"))
full = rcm.read_atm(os.path.join(G, "column21.atm"))
pl = atm[:, 1].copy()
Tlev, vlev = rcm.make_ensemble(ncol, 4242, pl, atm[:, 2].copy(), full[:, 4:9].T.copy())
st = rcm.init_columns(pl, Tlev, vlev)
h2o_ref, o3_ref = st["vmr9"][0, 0].copy(), st["vmr9"][0, 2].copy()
wvl, tau5 = rcm.make_lbl_tables(nwvl, 777, pl, h2o_ref, o3_ref)
s = rcm.Solver(0)
s.set_lbl_tables(wvl, tau5, h2o_ref, o3_ref, co2)
s.set_columns(pl, st["Tlayer"], Tlev[:, 20].copy(), st["vmr9"], st["rel_hum"])
s.advance(2)
s.synchronize()
t0 = time.perf_counter()
n = 5
s.advance(n)
s.synchronize()
dt = (time.perf_counter() - t0) / n
units = ncol * nwvl * 20
print(f"LBL ncol={ncol} nwvl={nwvl} co2x{co2}: {dt*1e3:.3f} ms/step, {units/dt/1e9:.2f} G updates/s, OLR[0]={s.get_state()['E_up'][0,0]:.6f}")
