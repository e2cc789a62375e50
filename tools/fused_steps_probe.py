"""Kernel time per step when n steps are fused into one launch (state resident in shared memory between the steps of a tile):
separates the per-launch / per-tile costs of the step kernel from its per-step cost.  Usage: python tools/fused_steps_probe.py [ncol]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench, our_first_climate_model_b200 as rcm
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
st = bench.build_ensemble(rcm, ncol, 12345)
s = rcm.Solver(0)
s.set_repwvl_table_from(rcm.Table(os.path.join(bench.GOLDEN, "Reduced100Forcing.rcmtab")))
s.set_columns(st["plevel"], st["Tlayer"], st["Tsurf"], st["vmr9"], st["rel_hum"])
s.advance(3, want_scalars=False)
print(f"{ncol} columns = {ncol / 16 / 444:.3f} rounds of 444 tiles")
for n in (1, 4, 16, 64):
    s.synchronize(); s.kernel_time_ms(reset=True)
    reps = max(2, 32 // n)
    for _ in range(reps):
        s.advance_async(n)
    s.synchronize()
    ms, k = s.kernel_time_ms(reset=True)
    print(f"{n:3d} steps per launch: {ms:8.3f} ms per launch = {ms / n:.4f} ms per step ({k} launches)")
