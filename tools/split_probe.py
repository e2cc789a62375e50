"""GPU probe: per-step device time of the split path and of the fused tile kernel at several ensemble sizes,
as one launch per step and as blocks of fused steps.  python tools/split_probe.py [ncol ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import our_first_climate_model_b200 as rcm  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [65536, 8192]
    atm = rcm.read_atm(os.path.join(G, "column21.atm"))
    pl = atm[:, 1].copy()
    for ncol in sizes:
        Tlev, vlev = rcm.make_ensemble(ncol, 12345, pl, atm[:, 2].copy(), atm[:, 4:9].T.copy())
        st = rcm.init_columns(pl, Tlev, vlev)
        for path, name in ((0, "split"), (1, "fused")):
            s = rcm.Solver(0)
            s.set_option(5, path)
            s.set_repwvl_table_from(rcm.Table(os.path.join(G, "Reduced100Forcing.rcmtab")))
            s.set_columns(pl, st["Tlayer"], np.full(ncol, 288.2), st["vmr9"], st["rel_hum"])
            s.advance(3)
            for block, reps in ((1, 30), (50, 2)):
                s.advance_async(block)  # buffers of this block size exist before the clock starts
                s.synchronize()
                s.kernel_time_ms(reset=True)
                t0 = time.perf_counter()
                for _ in range(reps):
                    s.advance_async(block)
                s.synchronize()
                wall = (time.perf_counter() - t0) / (reps * block) * 1e3
                kms, kn = s.kernel_time_ms(reset=True)
                per_step_k = kms if path == 0 else kms / block
                print(f"ncol {ncol:6d} {name} block {block:3d}: wall {wall:.4f} ms/step, dominant kernel {per_step_k:.4f} ms/step "
                      f"({kn} timed launches), {ncol * 2000 / wall / 1e6:.2f} G updates/s", flush=True)
            s.close()


if __name__ == "__main__":
    main()
