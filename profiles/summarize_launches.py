"""Summarise an ncu launch list (--metrics gpu__time_duration.sum[,smsp__inst_executed.sum,dram__bytes_read.sum] --csv)
per kernel: average duration, share of the step, warp instructions, DRAM bytes read."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:44]
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = row["Metric Unit"]
        metric = row["Metric Name"]
        if metric == "gpu__time_duration.sum":
            v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        elif metric == "dram__bytes_read.sum":
            v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
        agg[name][metric].append(v)
    tot = sum(sum(m["gpu__time_duration.sum"]) / len(m["gpu__time_duration.sum"]) for m in agg.values())
    print("%-46s %3s %9s %6s %9s %9s" % ("kernel", "n", "avg_us", "share", "winst_M", "dram_rd_MB"))
    for k, m in sorted(agg.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
        d = m["gpu__time_duration.sum"]
        avg = sum(d) / len(d)
        wi = m.get("smsp__inst_executed.sum")
        dr = m.get("dram__bytes_read.sum")
        print("%-46s %3d %9.2f %5.1f%% %9s %9s" % (k, len(d), avg, 100 * avg / tot,
                                                    "%.2f" % (sum(wi) / len(wi) / 1e6) if wi else "-",
                                                    "%.1f" % (sum(dr) / len(dr)) if dr else "-"))
    print("sum of kernel durations per step: %.1f us" % tot)


if __name__ == "__main__":
    main(sys.argv[1])
