#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (tools/calls/gpu_call20.sh) per kernel:
python tools/launch_summary.py gpurun_out/<tag>_launches.csv > profiles/<name>_summary.txt"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            ns = float(r["Metric Value"].replace(",", ""))
            if r.get("Metric Unit") in ("us", "usecond"):
                ns *= 1e3
            elif r.get("Metric Unit") in ("ms", "msecond"):
                ns *= 1e6
            name = re.sub(r"\(.*$", "", r["Kernel Name"])
            name = name.replace("beom::", "").replace("void ", "")
            rows.append((name, ns * 1e-6))
    agg = OrderedDict()
    for n, ms in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ms
    print("# kernel %-62s launches   total ms   mean ms" % "")
    for n, (k, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %6d %10.3f %9.4f" % (n[:70], k, t, t / k))
    fused = [ms for n, ms in rows if "k_fused_step" in n]
    print("# fused-step launches in order: " + " ".join("%.2f" % x for x in fused))
    tot = sum(ms for _, ms in rows)
    print("# all launches %.1f ms, k_fused_step %.1f ms (%.1f %%)" % (tot, sum(fused), 100.0 * sum(fused) / max(tot, 1e-9)))


if __name__ == "__main__":
    main(sys.argv[1])
