#!/usr/bin/env python
"""Condenses an `ncu -i X.ncu-rep --page raw --csv` dump into (1) a per-launch CSV of the metrics that decide what bounds
each kernel and (2) a JSON of per-kernel averages (duration, DRAM traffic per launch) that bench.py reads for
`roofline.traffic`.   usage: python tools/ncu_summary.py gpurun_out/stage_v8_raw.csv profiles/r01_ncu_full_v8"""
import csv
import json
import re
import sys

WANT = [
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue_slots_busy_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem_B"),
    ("launch__shared_mem_per_block_static", "static_smem_B"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp_latency_per_inst"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instruction"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch"),
]


def main():
    src, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, n) for m, n in WANT if m in idx]
    scale = {}
    for m, n in cols:
        u = units[idx[m]].lower()
        s = 1.0
        if n == "dur_us":
            s = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        if n.endswith("_MB"):
            s = {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1.0)
        scale[m] = s
    per = {}
    with open(out + "_summary.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [n for _, n in cols])
        for r in rows[2:]:
            name = re.sub(r"^void ", "", r[idx["Kernel Name"]]).split("(")[0]
            vals = []
            for m, n in cols:
                try:
                    v = float(r[idx[m]].replace(",", "")) * scale[m]
                except ValueError:
                    v = float("nan")
                vals.append(v)
            w.writerow([name] + ["%.6g" % v for v in vals])
            per.setdefault(name, []).append(dict(zip([n for _, n in cols], vals)))
    avg = {}
    for k, lst in per.items():
        a = {n: sum(d[n] for d in lst) / len(lst) for n in lst[0]}
        a["launches"] = len(lst)
        a["dram_bytes_per_launch"] = (a["dram_read_MB"] + a["dram_write_MB"]) * 1e6
        avg[k] = a
    with open(out + "_avg.json", "w") as f:
        json.dump({"source": src, "note": "ncu --set full --clock-control none, steady state of bench.py (NVTX range kernel_timing); "
                   "per-launch averages; DRAM traffic = dram__bytes_read.sum + dram__bytes_write.sum", "kernels": avg}, f, indent=1)
    for k, a in sorted(avg.items(), key=lambda kv: -kv[1]["dur_us"]):
        print("%-28s n=%2d  %7.1f us  dram %6.1f MB  issue %4.1f%%  sm %4.1f%%  l2hit %4.1f%%  occ %4.1f%%  regs %d" % (
            k, a["launches"], a["dur_us"], a["dram_bytes_per_launch"] / 1e6, a["issue_slots_busy_pct"], a["sm_pct"], a["l2_hit_pct"],
            a["achieved_occupancy_pct"], a["regs"]))


if __name__ == "__main__":
    main()
