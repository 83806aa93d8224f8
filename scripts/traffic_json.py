#!/usr/bin/env python
"""profiles/rNN_full_summary.csv (scripts/ncu_summary.py) -> profiles/rNN_traffic.json: DRAM bytes per launch per kernel, the file
bench.py reads for `roofline.traffic`.   usage: python scripts/traffic_json.py profiles/r02a_full_summary.csv profiles/r02a_traffic.json"""
import csv
import json
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hdr = rows[0]
    col = lambda key: next(i for i, h in enumerate(hdr) if h.startswith(key))
    unit = lambda key: re.search(r"\[(.*?)\]", hdr[col(key)]).group(1)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    t_scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    out = {}
    for r in rows[1:]:                             # same-named launches (ray samples / TV edge points, or two captured steps): keep the
        name = re.sub(r"^void ", "", r[0])         # LARGEST launch — it is the one the bench's per-step time is dominated by
        name = re.sub(r"\(.*", "", name).strip()
        rec = {"dram_bytes_per_launch": float(r[col("dram_read")]) * scale[unit("dram_read")] + float(r[col("dram_write")]) * scale[unit("dram_write")],
               "ms_per_launch": float(r[col("time")]) * t_scale[unit("time")]}
        for key in ("red_sectors", "ld_sectors"):
            try:
                rec[key + "_per_launch"] = float(r[col(key)])
            except StopIteration:
                pass
        prev = out.get(name)
        rec["launches"] = (prev["launches"] if prev else 0) + 1
        if prev is None or rec["ms_per_launch"] > prev["ms_per_launch"]:
            out[name] = rec
        else:
            prev["launches"] = rec["launches"]
    json.dump({"source": f"{src} (ncu --set full --clock-control none, bench.py --steps 1 --warmup 1)", "kernels": out}, open(dst, "w"), indent=1)
    print(f"{len(out)} kernels -> {dst}")


if __name__ == "__main__":
    main()
