"""Per-kernel summary of an `ncu --set full` capture exported with `ncu -i X.ncu-rep --page raw --csv` (gzipped or not):
python tools/ncu_per_kernel.py gpurun_out/r2_full_raw.csv.gz > profiles/r2_per_kernel.md"""
import collections
import csv
import gzip
import re
import sys

COLS = [
    ("gpu__time_duration.sum", "ms", 1.0),
    ("dram__bytes_read.sum", "MB read", 1.0),
    ("dram__bytes_write.sum", "MB written", 1.0),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1.0),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1.0),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1.0),
    ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
]


def num(v, unit, want):
    v = float(v.replace(",", "")) if v not in ("", "n/a") else 0.0
    u = unit.lower()
    if want == "ms":
        return v / 1e6 if u.startswith("ns") else v / 1e3 if u.startswith("us") else v * 1e3 if u in ("s", "second") else v
    if want.startswith("MB"):
        return v / 1e6 if u == "byte" else v / 1e3 if u == "kbyte" else v * 1e3 if u == "gbyte" else v
    return v


def main(path):
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.reader(op(path, "rt")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in data:
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").strip()
        a = agg.setdefault(name, {"n": 0, "vals": collections.defaultdict(float), "grid": r[ix["launch__grid_size"]],
                                  "block": r[ix["launch__block_size"]]})
        a["n"] += 1
        dur = num(r[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]], "ms")
        for col, label, _ in COLS:
            v = num(r[ix[col]], units[ix[col]], label)
            if label in ("ms", "MB read", "MB written"):
                a["vals"][label] += v
            elif label == "regs":
                a["vals"][label] = v
            else:
                a["vals"][label] += v * dur                # time-weighted mean of the percentages
        a["vals"]["_w"] += dur
    tot = sum(a["vals"]["ms"] for a in agg.values())
    labels = [c[1] for c in COLS]
    print("| kernel | launches | " + " | ".join(labels) + " | share of pass |")
    print("|---|---:|" + "---:|" * (len(labels) + 1))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["vals"]["ms"]):
        cells = []
        for label in labels:
            v = a["vals"][label]
            if label not in ("ms", "MB read", "MB written", "regs"):
                v = v / a["vals"]["_w"] if a["vals"]["_w"] else 0.0
            cells.append("%.3f" % v if label == "ms" else "%.1f" % v if label != "regs" else "%d" % v)
        print("| `%s` | %d | %s | %.1f%% |" % (name, a["n"], " | ".join(cells), 100 * a["vals"]["ms"] / tot))
    print("| **all** | %d | %.3f | | | | | | | | | | | 100%% |" % (sum(a["n"] for a in agg.values()), tot))


if __name__ == "__main__":
    main(sys.argv[1])
