"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table of ONE training step:
the launches after the second-to-last `adam_kernel` up to and including the last one.

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/<round>_launch_summary.txt
"""
import csv
import re
import sys


def main(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    launches = []
    for r in rows[1:]:
        us = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
        launches.append((re.sub(r"\(.*", "", r[ki]).strip(), us))
    adam = [i for i, (k, _) in enumerate(launches) if "adam_kernel" in k]
    if len(adam) >= 2:
        launches = launches[adam[-2] + 1:adam[-1] + 1]
    tot = sum(u for _, u in launches)
    agg = {}
    for k, u in launches:
        a = agg.setdefault(k, [0.0, 0])
        a[0] += u
        a[1] += 1
    print("ncu --metrics gpu__time_duration.sum --clock-control none over one eager training step "
          "(python bench.py --steps 1 --warmup 3 --no-graph):")
    print(f"{len(launches)} launches, {tot:.1f} us total (cold-cache, serialised: compare SHARES with bench.py's kernel table, not absolutes)\n")
    print("  total us   share  launches   avg us  kernel")
    for k, (u, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{u:10.1f} {100 * u / tot:6.1f}% {n:9d} {u / n:8.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv")
