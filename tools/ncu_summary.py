#!/usr/bin/env python
"""Selected metrics of an ncu report as a small CSV (metric, unit, value) for profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_full_x.csv
Reads the report with `ncu -i <rep> --page raw --csv` (first profiled launch)."""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(Kernel Name|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|sm__cycles_active\.avg|sm__cycles_elapsed\.max|"
    r"smsp__inst_executed\.sum|sm__inst_executed_pipe_(fp64|tensor.*|lsu|alu|fma.*|xu|uniform)\.sum|sm__ops_path_tensor_src_fp64\.(sum|avg\.pct_of_peak_sustained_elapsed|avg\.peak_sustained)|"
    r"sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_fp64\.avg\.pct_of_peak_sustained_active|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__issue_active\.avg\.per_cycle_active|smsp__warps_eligible\.avg\.per_cycle_active|smsp__warps_active\.avg\.per_cycle_active|"
    r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
    r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|launch__(block_size|grid_size|registers_per_thread|"
    r"shared_mem_per_block_dynamic|occupancy_limit_.*|waves_per_multiprocessor|sm_count)|sm__warps_active\.avg\.pct_of_peak_sustained_active)$")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, vals = rows[0], rows[1], rows[2]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for n, u, v in zip(names, units, vals):
            if KEEP.match(n):
                w.writerow([n, u, v])
    print(out)


if __name__ == "__main__":
    main()
