"""GPU probe: blocks of steps through the multi-step unit kernel (RCM_OPT_MULTI_STEP, option 6) against three launches per step:
time per step and bit-identity of the final state.  python tools/multi_probe.py [ncol ...]"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import our_first_climate_model_b200 as rcm  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [8192, 65536]
    atm = rcm.read_atm(os.path.join(G, "column21.atm"))
    pl = atm[:, 1].copy()
    for ncol in sizes:
        Tlev, vlev = rcm.make_ensemble(ncol, 12345, pl, atm[:, 2].copy(), atm[:, 4:9].T.copy())
        st = rcm.init_columns(pl, Tlev, vlev)
        for multi in (2, 0):
            s = rcm.Solver(0)
            s.set_option(6, multi)
            s.set_repwvl_table_from(rcm.Table(os.path.join(G, "Reduced100Forcing.rcmtab")))
            s.set_columns(pl, st["Tlayer"], np.full(ncol, 288.2), st["vmr9"], st["rel_hum"])
            sc = s.advance(3)
            for block, reps in ((50, 4), (250, 2)):
                s.advance_async(block)
                s.synchronize()
                best = 1e9
                for _ in range(2):
                    t0 = time.perf_counter()
                    for _ in range(reps):
                        s.advance_async(block)
                    s.synchronize()
                    best = min(best, (time.perf_counter() - t0) / (reps * block) * 1e3)
                print(f"ncol {ncol:6d} multi {multi} block {block:3d}: {best:.4f} ms/step", flush=True)
            sc = s.advance(7)
            g = s.get_state()
            h = hashlib.sha256(b"".join(g[k].tobytes() for k in ("Tlayer", "Tsurf", "E_up", "E_down", "dE", "h2o", "dt", "time_h")) + sc.tobytes())
            print(f"ncol {ncol:6d} multi {multi} state sha {h.hexdigest()[:16]}", flush=True)
            s.close()


if __name__ == "__main__":
    main()
