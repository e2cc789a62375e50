"""Generate tests/golden/ref_*.npz by running the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference and `make -C oracle`).  The fixtures pin the
plain-C port (oracle/rcm_oracle.c) and the CUDA path on boxes where the reference itself is
not available.  Inputs are stored next to the outputs, so nothing depends on RNG streams.
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import our_first_climate_model_b200 as rcm  # noqa: E402  (host-only helpers: atm reader, ensemble)
from oracle import refcpu as R  # noqa: E402
from oracle import port as P  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def main():
    assert R.available(), "build oracle/_ref first (make -C oracle)"
    atm = rcm.read_atm(os.path.join(G, "column21.atm"))
    pl, baseT, basev = atm[:, 1], atm[:, 2], atm[:, 4:9].T.copy()
    solar = R.solar()["solar_irr"]

    # ---- set A: 16 columns for the full pipeline ---------------------------------------------
    Tlev, vlev = rcm.make_ensemble(14, 12345, pl, baseT, basev)
    tab100 = P.load_rcmtab(os.path.join(G, "Reduced100Forcing.rcmtab"))
    st = R.init_columns(pl, Tlev, vlev)
    # two members whose layer temperatures sit exactly on table nodes (LowerPos exact-hit rule)
    _, lp, _ = P.read_tau(tab100, pl, st["Tlayer"][0], st["vmr9"][0])
    node_T = np.array([tab100["t_ref"][lp[19 - l]] + tab100["t_pert"][4 if l % 2 else 3] for l in range(20)])
    Tl_exact = np.stack([node_T, node_T[::-1].copy()])
    Tlayer = np.concatenate([st["Tlayer"], Tl_exact])
    vmr9 = np.concatenate([st["vmr9"], st["vmr9"][:2]])
    rel_hum = np.concatenate([st["rel_hum"], st["rel_hum"][:2]])
    Tsurf = np.full(16, 288.2)
    Tsurf[1:14] = Tlev[1:14, 20]
    out = dict(plevel=pl, Tlevel=Tlev, vmr_ppm_level=vlev, Tlayer=Tlayer, vmr9=vmr9, rel_hum=rel_hum, Tsurf=Tsurf,
               solar_irr=solar, init_Tlayer=st["Tlayer"], init_vmr9=st["vmr9"], init_rel_hum=st["rel_hum"],
               player=st["player"], conv=st["conv"])
    for n in (10, 20, 100):
        path = os.path.join(G, f"Reduced{n}Forcing.rcmtab")
        taus, lps, lts = [], [], []
        tab = P.load_rcmtab(path)
        for c in range(16):
            t, _, _ = R.read_tau(path, pl, Tlayer[c], vmr9[c], cloud_on=True)
            taus.append(t)
            _, lp, lt = P.read_tau(tab, pl, Tlayer[c], vmr9[c])
            # the index arrays come from the port; the reference's own LowerPos is checked below
            for k in range(20):
                midP = (pl[20 - k] * 100.0 + pl[19 - k] * 100.0) / 2
                assert R.lowerpos(tab["p_grid"], midP) == lp[k]
                tl = tab["t_ref"][lp[k]] + tab["t_pert"]
                assert R.lowerpos(tl, Tlayer[c][19 - k]) == lt[k]
            lps.append(lp)
            lts.append(lt)
        out[f"tau{n}"] = np.array(taus)
        out[f"lowpos_p{n}"] = np.array(lps, dtype=np.int32)
        out[f"lowpos_t{n}"] = np.array(lts, dtype=np.int32)
        r1 = R.advance(path, pl, rel_hum, solar, Tlayer, Tsurf, vmr9, 1)
        for k in ("E_down", "E_up", "dE", "dt", "Tlayer", "Tsurf", "time_h"):
            out[f"s1_{k}_{n}"] = r1[k]
        r5 = R.advance(path, pl, rel_hum, solar, Tlayer, Tsurf, vmr9, 5, want_trace=True)
        for k in ("E_down", "E_up", "dE", "dt", "Tlayer", "Tsurf", "time_h", "trace"):
            out[f"s5_{k}_{n}"] = r5[k]
        out[f"s5_h2o_{n}"] = r5["vmr9"][:, 0]
    # long run (Reduced100, first 4 columns): 300 steps, and the base column to stationarity
    path = os.path.join(G, "Reduced100Forcing.rcmtab")
    r = R.advance(path, pl, rel_hum[:4], solar, Tlayer[:4], Tsurf[:4], vmr9[:4], 300)
    out["s300_Tlayer_100"], out["s300_Tsurf_100"], out["s300_E_up_100"] = r["Tlayer"], r["Tsurf"], r["E_up"]
    r = R.advance(path, pl, rel_hum[:1], solar, Tlayer[:1], Tsurf[:1], vmr9[:1], 6000)
    out["s6000_Tlayer_100"], out["s6000_Tsurf_100"] = r["Tlayer"], r["Tsurf"]
    np.savez_compressed(os.path.join(G, "ref_repwvl.npz"), **out)

    # ---- set B: edge members, tau + indices only (out-of-range and exact hits) ----------------
    Te = np.repeat(st["Tlayer"][:1], 4, axis=0).copy()
    _, lp, _ = P.read_tau(tab100, pl, st["Tlayer"][0], st["vmr9"][0])
    for l in range(20):
        tref = tab100["t_ref"][lp[19 - l]]
        Te[0, l] = tref - 121.0          # below the lowest node: LowerPos falls through to the LAST interval
        Te[1, l] = tref + 125.0          # above the highest node
        Te[2, l] = tref + tab100["t_pert"][0] if l % 2 else tref + tab100["t_pert"][8]  # first / last node exactly
        Te[3, l] = tref + tab100["t_pert"][l % 9]
    ve = np.repeat(st["vmr9"][:1], 4, axis=0)
    edge = dict(Tlayer=Te, vmr9=ve, plevel=pl)
    for n in (10, 100):
        path = os.path.join(G, f"Reduced{n}Forcing.rcmtab")
        tab = P.load_rcmtab(path)
        taus, lts = [], []
        for c in range(4):
            t, _, _ = R.read_tau(path, pl, Te[c], ve[c], cloud_on=True)
            taus.append(t)
            lts.append([R.lowerpos(tab["t_ref"][lp[k]] + tab["t_pert"], Te[c][19 - k]) for k in range(20)])
        edge[f"tau{n}"] = np.array(taus)
        edge[f"lowpos_t{n}"] = np.array(lts, dtype=np.int32)
    edge["lowpos_p"] = np.array(lp, dtype=np.int32)
    np.savez_compressed(os.path.join(G, "ref_edge.npz"), **edge)

    # ---- LowerPos truth table and cplkavg samples ----------------------------------------------
    rng = np.random.default_rng(7)
    asc, desc = np.array([0.0, 1.0, 2.0, 3.0]), np.array([3.0, 2.0, 1.0, 0.0])
    xs = np.array([-1, 0, 0.5, 1, 1.5, 2, 2.5, 3, 4], dtype=float)
    lo = 10 ** rng.uniform(3.0, 5.3, 400)
    width = np.concatenate([10 ** rng.uniform(-6, -2.2, 200), 10 ** rng.uniform(-1.9, 0.7, 200)])
    hi = lo * (1 + width)
    T = rng.uniform(150, 350, 400)
    misc = dict(lp_x=xs, lp_asc=np.array([R.lowerpos(asc, x) for x in xs]),
                lp_desc=np.array([R.lowerpos(desc, x) for x in xs]), cpl_lo=lo, cpl_hi=hi, cpl_T=T,
                cpl_val=R.cplkavg_many(lo, hi, T))
    s = R.solar()
    misc.update({f"solar_{k}": v for k, v in s.items()})
    np.savez_compressed(os.path.join(G, "ref_misc.npz"), **misc)
    print("golden fixtures written to", G)


if __name__ == "__main__":
    main()
