"""Condense an `ncu --page raw --csv` export into one row per captured launch with the metrics DESIGN.md quotes.

    python tools/ncu_summary.py gpurun_out/r2f_ncu_full_raw.csv r2_end > profiles/r2_ncu_full_summary.csv
"""
import csv
import re
import sys

WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__block_size",
        "launch__grid_size", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(path, tag):
    rows = list(csv.reader(open(path, errors="replace")))
    head, units = rows[0], rows[1]
    col = {}
    for w in WANT:
        for i, h in enumerate(head):
            if h == w or h.endswith("." + w):
                col[w] = i
                break
    ki = head.index("Kernel Name")
    out = csv.writer(sys.stdout)
    out.writerow(["capture", "id", "kernel"] + [f"{w} [{units[col[w]]}]" for w in WANT if w in col])
    for r in rows[2:]:
        if len(r) <= ki:
            continue
        name = re.sub(r"^void vitb::", "", re.sub(r"\(.*", "", r[ki])).strip()
        out.writerow([tag, r[0], name] + [r[col[w]].replace(",", "") for w in WANT if w in col])


main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "capture")
