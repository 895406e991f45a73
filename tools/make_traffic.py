#!/usr/bin/env python
"""profiles/traffic.json from an ncu --set full summary (tools/ncu_summary.py) of ONE frame's kernels:
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per bench stage, per launch sequence of one
frame.  Usage: python tools/make_traffic.py summary.json profiles/traffic.json [more_summaries.json ...]"""
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(k, name):
    v = k.get(name)
    if isinstance(v, list):
        return float(v[0]) * UNIT.get(v[1], 1.0)
    return float(v or 0.0)


def main(src, dst, *more):
    ks = json.load(open(src))
    for m in more:
        ks += json.load(open(m))
    out = {}
    n_sweep = 0
    for k in ks:
        name = k["kernel"]
        b = val(k, "dram__bytes_read.sum") + val(k, "dram__bytes_write.sum")
        if "preprocess" in name:
            st = "preprocess"
        elif "onesweep" in name:
            st = "depth_sort" if n_sweep < 4 else "tile_sort"
            n_sweep += 1
        elif "hist_kernel" in name or "scan_rows" in name:  # compact_hist_kernel too
            st = "depth_sort"
        elif "emit" in name or "count_kernel" in name or "pair_scan" in name:
            st = "emit"
        elif "tile_scan" in name:
            st = "tile_scan"
        elif "ranges" in name or "tile_order" in name:
            st = "tile_sort"
        elif "composite" in name:
            st = "composite"
        elif "pose_kernel" in name:
            st = "pose"
        elif "pack_masks" in name:
            st = "pack_masks"
        elif "pack_kernel" in name:
            st = "pack_frame"
        else:
            continue
        out[st] = out.get(st, 0.0) + b
    out["_source"] = src
    json.dump(out, open(dst, "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main(*sys.argv[1:])
