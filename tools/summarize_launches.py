#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys

def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    agg, tot = collections.OrderedDict(), 0.0
    for r in csv.DictReader(lines):
        short = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "").replace("bb::", "")
        short = short.replace("Fe<FpCfg>", "Fp").replace("Fe<FrCfg>", "Fr")
        t = float(r["Metric Value"]) / 1e6
        a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
    print(f"# {path}: {tot:.2f} ms of kernel time over {sum(a[0] for a in agg.values())} launches (ncu-serialised, cold cache: compare shares)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:9.3f} ms  {100*t/tot:5.1f}%  n={n:4d}  avg {t/n:8.3f} ms  {k}")

if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
