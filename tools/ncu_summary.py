"""Condenses an `ncu --page raw --csv` export into the per-launch summary committed under profiles/ and refreshes
profiles/traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch, what bench.py reports as roofline.traffic).

    python tools/ncu_summary.py gpurun_out/r02_ncu_full_raw.csv profiles/r02_ncu_full_summary.csv [--traffic-tag "@2^28"]
"""
import csv
import json
import os
import re
import sys

COLS = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    src, dst = sys.argv[1], sys.argv[2]
    tag = sys.argv[sys.argv.index("--traffic-tag") + 1] if "--traffic-tag" in sys.argv else None
    rows = list(csv.reader(open(src)))
    head, units = rows[0], rows[1]
    idx = [head.index(c) for c in COLS if c in head]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([head[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    if tag:
        p = os.path.join(ROOT, "profiles", "traffic.json")
        tr = json.load(open(p)) if os.path.exists(p) else {}
        ki, ri, wi = head.index("Kernel Name"), head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        for r in rows[2:]:
            name = re.sub(r"^void ", "", r[ki]).split("<")[0].split("(")[0]
            if f"{name}{tag}" in tr and not sys.argv.count("--overwrite"):
                continue
            tr[f"{name}{tag}"] = float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
        tr["_source_" + os.path.basename(dst)] = f"{os.path.basename(dst)}: ncu --set full --clock-control none, one launch each"
        json.dump(tr, open(p, "w"), indent=1)
    print(f"{dst}: {len(rows) - 2} launches")


if __name__ == "__main__":
    main()
