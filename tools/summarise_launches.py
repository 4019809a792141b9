"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
python tools/summarise_launches.py profiles/r1_launches.csv > table.md"""
import collections
import csv
import re
import sys


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]).strip(), ms))
    tot = sum(ms for _, ms in rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, ms in rows:
        agg[k][0] += 1
        agg[k][1] += ms
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% |" % (k, n, ms, 100 * ms / tot))
    print("| **all** | %d | %.3f | 100%% |" % (len(rows), tot))


if __name__ == "__main__":
    main(sys.argv[1])
