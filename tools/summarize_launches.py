"""Summarise an ncu launch list (--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv):
per kernel name -> launches, total / average duration, share of the captured time, DRAM bytes per launch.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launch_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    per = defaultdict(lambda: defaultdict(float))  # launch id -> metric -> value
    name = {}
    for r in rd:
        i = int(r["ID"])
        name[i] = r["Kernel Name"]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        m = r["Metric Name"]
        if m.startswith("gpu__time_duration"):
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)  # -> us
        elif m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        per[i][m] = v
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i, m in per.items():
        n = re.sub(r"\(.*", "", name[i])
        n = re.sub(r"^void ", "", n)
        if "spin_kernel" in n:  # torch.cuda._sleep: bench.py holds the stream while it queues the roofline step
            continue
        a = agg[n]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0)
        a[3] += m.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print(f"{sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time captured (ncu per-launch times: cold caches, serialised)")
    print(f"{'share':>6s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'rd MB/launch':>12s} {'wr MB/launch':>12s}  kernel")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * a[1] / tot:5.1f}% {a[0]:8d} {a[1]:10.1f} {a[1] / a[0]:8.1f} {a[2] / a[0] / 1e6:12.2f} {a[3] / a[0] / 1e6:12.2f}  {n}")
    fam = [a for n, a in agg.items() if "gemm" in n]
    if fam:
        L = sum(a[0] for a in fam)
        print(f"tcgen05 GEMM family: {L} launches, {100 * sum(a[1] for a in fam) / tot:.1f}% of the captured time, "
              f"DRAM traffic {sum(a[2] + a[3] for a in fam) / L / 1e6:.1f} MB per launch "
              f"(read {sum(a[2] for a in fam) / L / 1e6:.1f} + write {sum(a[3] for a in fam) / L / 1e6:.1f})")


if __name__ == "__main__":
    main(sys.argv[1])
