#!/usr/bin/env python
"""Aggregate an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into per-kernel shares.
usage: python scripts/launch_share.py gpurun_out/launches.csv profiles/r01_launch_share.md"""
import collections
import csv
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        n = re.sub(r"\(.*", "", r[ki])[:80]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for n, a in agg.items() if "f2b::" in n)
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\nserialised, cold-cache per-launch times; {len(rows) - 1} launches, "
                f"{tot / 1e6:.3f} ms total, f2b kernels {ours / tot * 100:.1f} % of it.\n\n| share | launches | total us | kernel |\n|---|---|---|---|\n")
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {a[1] / tot * 100:.2f} % | {a[0]} | {a[1] / 1e3:.1f} | `{n}` |\n")
    print(f"{len(agg)} kernels -> {dst}")


if __name__ == "__main__":
    main()
