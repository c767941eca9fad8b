"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name."""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "ms": 1e3}.get(unit, 1e-3)
        tot[name] += v * scale
        cnt[name] += 1
    total = sum(tot.values())
    print("total kernel time %.3f ms over %d launches (%s)" % (total / 1e3, sum(cnt.values()), path))
    print("%10s %6s %7s  %s" % ("us", "n", "share", "kernel"))
    for name, t in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
        print("%10.1f %6d %6.1f%%  %s" % (t, cnt[name], 100 * t / total, name[:110]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
