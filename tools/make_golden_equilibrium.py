"""Generate tests/golden/ref_equilibrium.npz: the 16 columns of ref_repwvl.npz (member 0 unperturbed, 13 PERTURBED
ensemble members, 2 members sitting exactly on table nodes) taken through 6,000 iterations of the reference's time loop
(main.cpp:531-583) by the UNMODIFIED reference (oracle/_ref), one process per column.  Pins the north-star criterion
"equilibrium temperature profile within 1e-3 K" for perturbed members, not just for the base column.
Run in the build container (needs /root/reference and `make -C oracle`); ~1 minute on 8 cores."""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
G = os.path.join(ROOT, "tests", "golden")
NSTEPS = 6000


def one(c):
    from oracle import refcpu as R
    g = np.load(os.path.join(G, "ref_repwvl.npz"))
    r = R.advance(os.path.join(G, "Reduced100Forcing.rcmtab"), g["plevel"], g["rel_hum"][c:c + 1], float(g["solar_irr"]),
                  g["Tlayer"][c:c + 1], g["Tsurf"][c:c + 1], g["vmr9"][c:c + 1], NSTEPS)
    return r["Tlayer"][0], r["Tsurf"][0], r["E_up"][0], r["dt"][0], r["vmr9"][0, 0]


def main():
    from oracle import refcpu as R
    assert R.available(), "build oracle/_ref first (make -C oracle)"
    with mp.get_context("fork").Pool(min(16, os.cpu_count() or 1)) as pool:
        res = pool.map(one, range(16))
    out = dict(nsteps=NSTEPS, Tlayer=np.array([r[0] for r in res]), Tsurf=np.array([r[1] for r in res]),
               E_up=np.array([r[2] for r in res]), dt=np.array([r[3] for r in res]), h2o=np.array([r[4] for r in res]))
    g = np.load(os.path.join(G, "ref_repwvl.npz"))
    assert np.array_equal(out["Tlayer"][:1], g["s6000_Tlayer_100"])  # same run as the one already pinned for member 0
    np.savez_compressed(os.path.join(G, "ref_equilibrium.npz"), **out)
    print("written", os.path.join(G, "ref_equilibrium.npz"), "Tsurf", out["Tsurf"])


if __name__ == "__main__":
    main()
