"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
and share of the device time.  Only the launches of the LAST forward are kept when --last-step
markers are not available; here we simply aggregate everything after the warm-up by kernel name."""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "")
    return name[:90]


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((short(r["Kernel Name"]), val * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for n, us in rows:
        agg[n][0] += 1
        agg[n][1] += us
    total = sum(v[1] for v in agg.values())
    print("launches: %d   total device time: %.1f us" % (len(rows), total))
    print("%-92s %6s %10s %7s %9s" % ("kernel", "count", "total_us", "share", "avg_us"))
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-92s %6d %10.1f %6.1f%% %9.2f" % (n, c, us, 100 * us / total, us / c))


if __name__ == "__main__":
    main(sys.argv[1])
