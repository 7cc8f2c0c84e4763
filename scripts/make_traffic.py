#!/usr/bin/env python
"""profiles/traffic.json from `ncu --set full` reports: DRAM bytes per launch of the dominant kernel
of each benchmarked configuration, stamped with the hash of the kernel sources the capture was
taken from (bench.py nulls `roofline.traffic` when the sources have changed since).

Usage: python scripts/make_traffic.py <tag> WORKLOAD:BATCH:XDT:YDT:report.ncu-rep ...
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402


def first_kernel(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    col = {h: i for i, h in enumerate(rows[0])}
    recs = []
    for vals in rows[2:]:
        recs.append({"kernel": vals[col["Kernel Name"]],
                     "bytes": float(vals[col["dram__bytes_read.sum"]].replace(",", "")) * _scale(rows[1][col["dram__bytes_read.sum"]])
                     + float(vals[col["dram__bytes_write.sum"]].replace(",", "")) * _scale(rows[1][col["dram__bytes_write.sum"]]),
                     "ms": float(vals[col["gpu__time_duration.sum"]].replace(",", "")) * _tscale(rows[1][col["gpu__time_duration.sum"]])})
    return recs


def _scale(unit):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def _tscale(unit):
    return {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1)


def main():
    tag = sys.argv[1]
    caps = []
    for spec in sys.argv[2:]:
        wl, batch, xdt, ydt, rep = spec.split(":", 4)
        recs = first_kernel(rep)
        # one launch = every kernel the apply issues (the compact path has two)
        names = []
        for r in recs:
            if r["kernel"] not in names:
                names.append(r["kernel"])
        per = {n: [r for r in recs if r["kernel"] == n] for n in names}
        total = sum(sum(r["bytes"] for r in v) / len(v) for v in per.values())
        caps.append({"workload": wl, "batch_rows": int(batch), "x_dtype": xdt, "y_dtype": ydt,
                     "kernel": " + ".join(names), "dram_bytes_per_launch": int(total),
                     "ms_under_ncu": round(sum(sum(r["ms"] for r in v) / len(v) for v in per.values()), 4),
                     "source": os.path.relpath(rep, ROOT)})
    doc = {"note": "dram__bytes_read.sum + dram__bytes_write.sum per apply, from `ncu --set full --clock-control none` "
                   "captures of `python bench.py --workload <w> --steps 3 --warmup 3 --no-cpu --no-e2e --no-others --strong-rows 0`",
           "round": tag, "kernel_source_hash": kernel_source_hash(), "captures": caps}
    json.dump(doc, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
