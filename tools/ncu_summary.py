"""Print the metrics that matter from an .ncu-rep (raw page csv) for every profiled launch."""
import argparse
import csv
import json
import os
import re
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum", "dram__sectors_read.sum", "lts__t_sectors_op_read.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
]


def _num(text):
    try:
        return float(text.replace(",", ""))
    except ValueError:
        return 0.0


def _bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return _num(value) * scale.get(unit, 1.0)


def main(path, traffic_json=None, workload=None, kernel_regex=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    best = None
    for r in rows[2:]:
        print("==== ", r[ki][:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:82s} {r[i]:>16s} {units[i]}")
        if traffic_json and re.search(kernel_regex, r[ki]):
            ir, iw, it = (hdr.index(k) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum",
                                                   "gpu__time_duration.sum"))
            rec = {"kernel": r[ki][:160], "dram_bytes_read": _bytes(r[ir], units[ir]),
                   "dram_bytes_write": _bytes(r[iw], units[iw]), "duration": _num(r[it]), "duration_unit": units[it],
                   "source": os.path.basename(path)}
            rec["dram_bytes_per_launch"] = rec["dram_bytes_read"] + rec["dram_bytes_write"]
            if best is None or rec["duration"] > best["duration"]:
                best = rec
    if traffic_json and best:
        # bench.py reads roofline.traffic from this file: DRAM bytes of one launch of the dominant kernel
        try:
            table = json.load(open(traffic_json))
        except (OSError, ValueError):
            table = {}
        table[workload] = best
        json.dump(table, open(traffic_json, "w"), indent=1, sort_keys=True)
        print(f"traffic of {workload}: {best['dram_bytes_per_launch']:.4g} bytes per launch -> {traffic_json}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--traffic-json", default=None)
    ap.add_argument("--workload", default=None)
    ap.add_argument("--kernel-regex", default=r"sb_fft_strided(32)?_kernel<.*?, *\(?int\)?1,")
    a = ap.parse_args()
    main(a.report, a.traffic_json, a.workload, a.kernel_regex)
