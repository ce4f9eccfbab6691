#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections, csv, re, sys

def main(path, top=30):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    det = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if "gpu__time_duration" not in row.get("Metric Name", ""):
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"mmfn::|tc::|<unnamed>::|\(anonymous namespace\)::|void ", "", name)
        t = float(row["Metric Value"].replace(",", "")) / 1e3
        agg[name][0] += 1; agg[name][1] += t; tot += t
        k = (name, row["Grid Size"]); det[k][0] += 1; det[k][1] += t
    print(f"launches {sum(v[0] for v in agg.values())}  total {tot/1e3:.2f} ms (serialised, cold cache: compare SHARES)")
    print("--- by kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]/1e3:8.3f} ms {100*v[1]/tot:5.1f}%  n={v[0]:4d} avg={v[1]/v[0]:7.1f} us  {k[:100]}")
    print("--- by kernel x grid")
    for k, v in sorted(det.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]/1e3:8.3f} ms n={v[0]:4d} avg={v[1]/v[0]:7.1f} us grid={k[1]:>16s}  {k[0][:80]}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
