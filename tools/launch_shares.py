"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total / median
duration, share of the listed launches.      python tools/launch_shares.py gpurun_out/launches.csv [out.json]"""
import collections
import csv
import json
import re
import statistics
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1e-3)
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", r[ki])
        name = re.sub(r"\(.*", "", name)
        per[name].append(v * scale)
    tot = sum(sum(v) for v in per.values())
    out = {k: {"launches": len(v), "total_us": round(sum(v), 1), "median_us": round(statistics.median(v), 2),
               "share": round(sum(v) / tot, 4)} for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1]))}
    for k, v in out.items():
        print(f"{k[:60]:60s} {v['launches']:5d} {v['total_us']:12.1f} us  med {v['median_us']:9.2f}  {100 * v['share']:5.1f} %")
    if len(sys.argv) > 2:
        json.dump({"total_us": round(tot, 1), "kernels": out}, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
