"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: sequence of the last step + per-kernel totals.

  python tools/summarize_launches.py gpurun_out/launches.csv [marker_kernel_substring] > summary.txt
The last occurrence of the marker kernel (default: prep_image) starts the step that is listed.
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("mv::", "")
    return name[:70]


def main():
    path = sys.argv[1]
    marker = sys.argv[2] if len(sys.argv) > 2 else "prep_image"
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        t = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") == "ns":
            t /= 1e3
        elif r.get("Metric Unit") == "ms":
            t *= 1e3
        rows.append((r["Kernel Name"], r["Grid Size"], r["Block Size"], t))
    marks = [i for i, r in enumerate(rows) if marker in r[0]]
    start, end = (marks[-1] if marks else 0), len(rows)
    # a launch-count limit (ncu -c N) may cut the last step short: then list the last COMPLETE one
    if len(marks) >= 3 and end - marks[-1] < marks[-1] - marks[-2]:
        start, end = marks[-2], marks[-1]
    step = rows[start:end]
    tot = sum(r[3] for r in step)
    print("# step: %d launches, %.1f us total (cold-cache, serialised ncu times)" % (len(step), tot))
    agg = OrderedDict()
    for n, g, b, t in step:
        k = short(n) + " " + g + " " + b
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    print("# per kernel (count, total us, share)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-110s %5d %10.1f %5.1f%%" % (k, c, t, 100 * t / tot))
    print("# sequence")
    for n, g, b, t in step:
        print("%-80s %-16s %-14s %9.1f" % (short(n), g, b, t))


if __name__ == "__main__":
    main()
