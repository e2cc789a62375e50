"""Clocks and power while the step kernel runs for a few seconds (is the FP64 load power-capped?)."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import our_first_climate_model_b200 as rcm
import bench
ncol = 65536
st = bench.build_ensemble(rcm, ncol, 1)
s = rcm.Solver(0)
s.set_repwvl_table_from(rcm.Table(os.path.join(bench.GOLDEN, "Reduced100Forcing.rcmtab")))
s.set_columns(st["plevel"], st["Tlayer"], st["Tsurf"], st["vmr9"], st["rel_hum"])
s.advance(3)
rows = []
p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu",
                      "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.strip()) for l in p.stdout], daemon=True).start()
time.sleep(0.3)
t0 = time.perf_counter()
for k in range(6):
    s.kernel_time_ms(reset=True)
    s.advance(100, want_scalars=False)
    s.synchronize()
    ms, n = s.kernel_time_ms(reset=True)
    print(f"block {k}: {ms/100:.3f} ms/step   t={time.perf_counter()-t0:.1f}s", flush=True)
time.sleep(0.2)
p.terminate()
print("idle->load samples (MHz, W, power_cap, C):")
for r in rows[::4]:
    print("   ", r)
