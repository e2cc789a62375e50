"""Quick GPU probe: FP64 microbenchmarks + timing of the fused step (development aid)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import our_first_climate_model_b200 as rcm

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 100
s = rcm.Solver(0)
names = ["dfma", "exp()", "div", "exp_tab"]
for w in range(4):
    print(f"microbench {names[w]:8s}: {s.fp64_microbench(w):10.1f} Gop/s", flush=True)
atm = rcm.read_atm(os.path.join(G, "column21.atm"))
pl = atm[:, 1]
Tlev, vlev = rcm.make_ensemble(ncol, 12345, pl, atm[:, 2], atm[:, 4:9].T.copy())
st0 = rcm.init_columns(pl, Tlev, vlev)
s.set_repwvl_table_from(rcm.Table(os.path.join(G, f"Reduced{nw}Forcing.rcmtab")))
cfgs = [tuple(int(x) for x in a.split(',')) for a in sys.argv[3:]] or [(1, 1), (1, 0)]
for cfg, pf in cfgs:
    s.set_option(1, cfg); s.set_option(3, pf)
    s.set_columns(pl, st0["Tlayer"], np.full(ncol, 288.2), st0["vmr9"], st0["rel_hum"])
    s.advance(2)
    s.kernel_time_ms(reset=True)
    t0 = time.time()
    s.advance(1); s.advance(1); s.advance(1)
    s.synchronize()
    wall = (time.time() - t0) / 3
    ms, n = s.kernel_time_ms(reset=True)
    units = ncol * nw * 20
    olr = s.get_state()["E_up"][0, 0]
    print(f"shape={cfg} stage_rows={pf} ncol={ncol} nwvl={nw}: kernel {ms:.3f} ms/step ({n} launches), wall {wall*1e3:.3f} ms/step, "
          f"{units/ms/1e6:.2f} Gunits/s  OLR[0]={olr:.10f}", flush=True)
    s.kernel_time_ms(reset=True)
    s.advance(10, want_scalars=False); s.synchronize()
    ms, n = s.kernel_time_ms(reset=True)
    print(f"   fused 10 steps in one launch: {ms/10:.3f} ms/step", flush=True)
