#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and
share of ONE sweep (the launches between two consecutive hamming_ll_block_kernel launches).
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [sweep_index_from_end=2]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    back = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [x["Kernel Name"] for x in rows]
    idx = [i for i, nm in enumerate(names) if "hamming_ll_block" in nm]
    a, b = idx[-back - 1], idx[-back]
    agg = collections.OrderedDict()
    tot = 0.0
    for x in rows[a:b]:
        nm = re.sub(r"\(.*", "", x["Kernel Name"])
        t = float(x["Metric Value"]) / 1000.0
        agg.setdefault(nm, [0, 0.0])
        agg[nm][0] += 1
        agg[nm][1] += t
        tot += t
    print(f"launches in the sweep: {b - a}; sum of kernel durations: {tot:.1f} us (cold-cache, serialised under ncu)")
    print(f"| kernel | launches | sum us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} | {v[1] / tot:.3f} |")


if __name__ == "__main__":
    main()
