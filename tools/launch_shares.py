#!/usr/bin/env python3
"""Per-kernel time shares of ONE V-cycle from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_shares.py gpurun_out/launches.csv [launches_per_cycle] > profiles/rNN_vcycle_kernel_shares.json

The last `launches_per_cycle` V-cycle kernels of the list are one eager (un-graphed under ncu) cycle; the
times are cold-cache and serialised, so only the SHARES are meaningful (see B200_PROFILING.md).
"""
import csv
import json
import re
import sys


def main():
    path = sys.argv[1]
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("cup::", "").replace("void ", "").strip()
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        rows.append((name, ns))
    mg = [(n, t) for n, t in rows if n.startswith(("k_smooth", "k_down", "k_up", "k_apply", "k_bottom"))]
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 43
    cyc = mg[-per:]
    tot = sum(t for _, t in cyc)
    k = {}
    for n, t in cyc:
        e = k.setdefault(n, {"us": 0.0, "launches": 0})
        e["us"] += t / 1e3
        e["launches"] += 1
    for e in k.values():
        e["share"] = round(e["us"] * 1e3 / tot, 4)
        e["us"] = round(e["us"], 1)
    out = {"source": path, "launches_per_cycle": per, "serialized_us": round(tot / 1e3, 1),
           "kernels": dict(sorted(k.items(), key=lambda kv: -kv[1]["us"]))}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
