#!/usr/bin/env python3
"""Text Gantt chart of one prove from a bench.py JSON line (key `timeline`): per MSM job the device
phases (sort = queued..accumulate start, accumulate, reduce = accumulate end..done) and the host
milestones, all in ms since bb_groth16_prove was entered.

    python bench.py > line.json;  python tools/timeline_report.py line.json [--width 100]
"""
import json
import sys


def main():
    path = sys.argv[1]
    width = int(sys.argv[sys.argv.index("--width") + 1]) if "--width" in sys.argv else 100
    line = [l for l in open(path).read().splitlines() if l.strip().startswith("{")][-1]
    d = json.loads(line)
    tl = d["timeline"]
    dev, host = tl["device_ms_since_prove_start"], tl.get("host_ms_since_prove_start", {})
    end = max([v["done"] for v in dev.values() if v["done"] is not None] + list(host.values()) + [d.get("ms_per_step", 0)])
    scale = width / end if end else 1.0

    def bar(segments):
        row = [" "] * (width + 1)
        for a, b, ch in segments:
            if a is None or b is None:
                continue
            for x in range(int(a * scale), max(int(a * scale) + 1, int(b * scale))):
                if x <= width:
                    row[x] = ch
        return "".join(row)

    print(f"{d['config']['workload']}: {d['ms_per_step']:.2f} ms per step ({d['n_gpus']} GPU); one column = {end / width:.2f} ms")
    print(f"{'job':12s} {'queued':>7s} {'acc0':>7s} {'acc1':>7s} {'done':>7s}  . sort  # accumulate  = reduce")
    for job, v in sorted(dev.items(), key=lambda kv: kv[1]["queued"] or 0):
        q, a0, a1, dn = v["queued"], v["accumulate_start"], v["accumulate_end"], v["done"]
        print(f"{job:12s} {q:7.2f} {a0:7.2f} {a1:7.2f} {dn:7.2f}  " + bar([(q, a0, "."), (a0, a1, "#"), (a1, dn, "=")]))
    if host:
        marks = sorted(host.items(), key=lambda kv: kv[1])
        row = [" "] * (width + 1)
        for i, (name, t) in enumerate(marks):
            row[min(width, int(t * scale))] = str(i + 1)
        print(f"{'host':12s} {'':31s}  " + "".join(row))
        print("   " + ", ".join(f"{i + 1} = {name} @ {t:.2f} ms" for i, (name, t) in enumerate(marks)))
        last_dev = max(v["done"] for v in dev.values() if v["done"] is not None)
        print(f"   device idle before the last kernel chain ends: see gaps above; host tail after the last device result: "
              f"{host.get('proof_done', 0) - last_dev:.2f} ms")


if __name__ == "__main__":
    main()
