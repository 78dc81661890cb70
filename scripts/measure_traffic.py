#!/usr/bin/env python
"""DRAM traffic per launch of the two lin_solve kernels, measured with ncu on the GPU box, written to
profiles/lin_solve_traffic.json together with the hash of the kernel sources that were profiled.  bench.py quotes the
record as `roofline.traffic` only while that hash matches the sources it runs (a stale record reads as null).

    gpurun -- python scripts/measure_traffic.py        (one GPU; ~1 minute)
"""
import csv
import datetime
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import source_sha  # noqa: E402

METRICS = "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"


def ncu_one(kernel_regex, size, k, orient, mode, skip):
    cmd = ["ncu", "--metrics", METRICS, "--clock-control", "none", "-k", f"regex:{kernel_regex}", "-s", str(skip), "-c", "1",
           "--csv", sys.executable, os.path.join(ROOT, "scripts", "prof_linsolve.py"), str(size), str(k), str(orient), "1", mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    rows = [r for r in csv.reader(io.StringIO(out.stdout)) if len(r) > 5]
    hdr = next(r for r in rows if "Metric Name" in r)
    iname, ival, iunit = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    vals = {}
    for r in rows:
        if r is hdr or len(r) <= ival:
            continue
        try:
            v = float(r[ival].replace(",", ""))
        except ValueError:
            continue
        unit = r[iunit]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1)
        vals[r[iname]] = v * scale
    return vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"], vals["gpu__time_duration.sum"]


def main():
    rec = {"source_sha": source_sha(), "when": datetime.datetime.utcnow().strftime("%Y-%m-%dT%H:%MZ"),
           "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one launch of "
                  "scripts/prof_linsolve.py (Passive solve on the config's grid and rectangles), cold cache"}
    for key, size, k in (("c4", 16384, 20), ("c3", 4096, 40)):
        b, ms = ncu_one("k_linsolve_tb", size, k, 2, "exact", 1)            # launch 0 is the K=1 warm-up
        rb, rms = ncu_one("k_rb_stream", size, k, 2, "red_black", 1)      # (launch 0: the first pass of the first solve)
        rec[key] = {"dram_bytes_per_launch": b, "launch_ms_under_ncu": ms, "rb_dram_bytes_per_launch": rb,
                    "rb_launch_ms_under_ncu": rms, "algorithmic_bytes_per_launch": 12.0 * (size - 2) ** 2 * k,
                    "rb_algorithmic_bytes_per_launch": 12.0 * (size - 2) ** 2 * 4}
        print(key, rec[key], flush=True)
    with open(os.path.join(ROOT, "profiles", "lin_solve_traffic.json"), "w") as fh:
        json.dump(rec, fh, indent=1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "lin_solve_traffic.json"), "w") as fh:
        json.dump(rec, fh, indent=1)


if __name__ == "__main__":
    main()
