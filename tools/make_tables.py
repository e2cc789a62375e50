"""Convert the reference's NetCDF-4 lookup tables into flat `.rcmtab` fixtures.

Run in the build container (needs /root/reference); the outputs are committed
under tests/golden/ so the GPU box (which has no /root/reference) can run the
parity tests and the bench.  Provenance: repwvl_V2.01_cpp/Reduced*Forcing.nc
(schema: SURVEY.md Appendix B), read with tools/nc4lite.py.

.rcmtab layout (little endian):
    char   magic[8]  = "RCMTAB01"
    u64    n_tpert, n_species, n_wvl, n_p
    f64    xsec[n_tpert][n_species][n_wvl][n_p]      (reference order: books, pages, rows, cols)
    f64    wvl[n_wvl]  weight[n_wvl]  p_grid[n_p]  t_ref[n_p]  t_pert[n_tpert]
    f64    vmrs_ref[n_species][n_p]
Also rewrites the two 21-level atmosphere files as whitespace-normalised
fixtures (4 header lines, then 21 rows; same numbers).
"""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
from nc4lite import read_nc  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")


def write_rcmtab(path, v):
    xs = v["xsec"]
    nb, npg, nr, nc = xs.shape
    assert v["ChosenWvls"].shape == (nr,) and v["ChosenWeights"].shape == (nr,)
    assert v["p_grid"].shape == (nc,) and v["t_ref"].shape == (nc,) and v["t_pert"].shape == (nb,)
    assert v["vmrs_ref"].shape == (npg, nc)
    with open(path, "wb") as f:
        f.write(b"RCMTAB01")
        f.write(struct.pack("<4Q", nb, npg, nr, nc))
        for k in ("xsec", "ChosenWvls", "ChosenWeights", "p_grid", "t_ref", "t_pert", "vmrs_ref"):
            f.write(np.ascontiguousarray(v[k], dtype="<f8").tobytes())


def rewrite_atm(src, dst, ncols):
    rows = []
    with open(src) as f:
        lines = f.read().splitlines()
    for ln in lines[4:]:
        tok = ln.split()
        if len(tok) == ncols:
            rows.append(tok)
    assert len(rows) == 21
    names = ["z[km]", "p[hPa]", "T[K]", "air[1/cm3]", "H2O[ppm]", "O3[ppm]", "CO2[ppm]", "CH4[ppm]", "N2O[ppm]"][:ncols]
    with open(dst, "w") as f:
        f.write("# 21-level column fixture (same numbers as the reference atmosphere file)\n#\n")
        f.write("# " + " ".join(names) + "\n#\n")
        for r in rows:
            f.write(" ".join(r) + "\n")


def main():
    os.makedirs(OUT, exist_ok=True)
    for n in (10, 20, 100):
        v = read_nc(f"{REF}/repwvl_V2.01_cpp/Reduced{n}Forcing.nc")
        write_rcmtab(os.path.join(OUT, f"Reduced{n}Forcing.rcmtab"), v)
        print("wrote table", n, v["xsec"].shape)
    rewrite_atm(f"{REF}/repwvl_V2.01_cpp/test.atm", os.path.join(OUT, "column21.atm"), 9)
    rewrite_atm(f"{REF}/lbl.arts/fpda.lbl.atm", os.path.join(OUT, "column21.lbl.atm"), 6)


if __name__ == "__main__":
    main()
