#!/usr/bin/env python
"""Write profiles/roofline_capture.json - the constants bench.py's roofline is computed from - out of ncu captures.

    tools/update_roofline_capture.py --tag r2c \
        --step-src gpurun_out/r2c_step_src.csv --step-raw gpurun_out/r2c_step_raw.csv --step-ncol 65536 --step-nwvl 100 \
        [--lbl-src ... --lbl-raw ... --lbl-ncol 512 --lbl-nwvl 20000]

*_src.csv: `ncu -i X.ncu-rep --page source --csv` of ONE launch of the kernel (per-SASS-line "Instructions Executed");
*_raw.csv: `ncu -i X.ncu-rep --page raw --csv` of the same launch (dram__bytes_read/write.sum, gpu__time_duration.sum).
FP64-pipe thread-instructions per unit = 32 x warp-instructions with a DFMA/DADD/DMUL/DSETP/DMNMX opcode / (ncol*nwvl*20).
The sha256 of the kernel sources (bench.KERNEL_SOURCES) is recorded with the numbers: bench.py prints "stale": true and
tests/test_host.py fails when the sources move without a new capture.  A workload that is not given keeps its old entry."""
import argparse
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")


def fold_source(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    col = {n: i for i, n in enumerate(rows[hi])}
    tot = fp = 0
    for r in rows[hi + 1:]:
        if len(r) < len(rows[hi]):
            continue
        toks = r[col["Source"]].split()
        op = toks[0] if toks and not toks[0].startswith("@") else (toks[1] if len(toks) > 1 else "?")
        e = int(r[col["Instructions Executed"]] or 0)
        tot += e
        if op.split(".")[0] in FP64:
            fp += e
    return tot, fp


def raw_metrics(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    out = {}
    for h, u, v in zip(hdr, units, vals):
        if h in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
            out[h] = float(v) * scale[u]
        if h == "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active":
            out["fp64_pipe_pct_active"] = float(v)
        if h == "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed":
            out["fp64_pipe_pct_elapsed"] = float(v)
        if h == "Kernel Name":
            out["kernel"] = v
    return out


def entry(tag, src, raw, ncol, nwvl):
    tot, fp = fold_source(src)
    m = raw_metrics(raw)
    units = ncol * nwvl * 20
    return {"capture": tag, "kernel": m.get("kernel", "?"), "ncol": ncol, "nwvl": nwvl,
            "warp_instructions": tot, "fp64_warp_instructions": fp, "exec_fp64_per_unit": 32.0 * fp / units,
            "dram_bytes_per_launch": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
            "dram_bytes_per_column_step": (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / ncol,
            "kernel_ms_under_ncu": m["gpu__time_duration.sum"],
            "fp64_pipe_pct_active": m.get("fp64_pipe_pct_active"), "fp64_pipe_pct_elapsed": m.get("fp64_pipe_pct_elapsed")}


def main():
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    for w in ("step", "lbl"):
        ap.add_argument(f"--{w}-src")
        ap.add_argument(f"--{w}-raw")
        ap.add_argument(f"--{w}-ncol", type=int)
        ap.add_argument(f"--{w}-nwvl", type=int)
    ap.add_argument("--keep-sha", action="store_true", help="leave sources_sha256 alone (a capture of an older build)")
    a = ap.parse_args()
    cap = json.load(open(bench.CAPTURE_JSON)) if os.path.exists(bench.CAPTURE_JSON) else {}
    for w in ("step", "lbl"):
        src, raw = getattr(a, f"{w}_src"), getattr(a, f"{w}_raw")
        if src and raw:
            cap[w] = entry(a.tag, src, raw, getattr(a, f"{w}_ncol"), getattr(a, f"{w}_nwvl"))
    if not a.keep_sha:
        cap["sources_sha256"] = bench.kernel_sources_sha()
        cap["sources"] = bench.KERNEL_SOURCES
    json.dump(cap, open(bench.CAPTURE_JSON, "w"), indent=1)
    print(json.dumps(cap, indent=1))


if __name__ == "__main__":
    main()
