#!/usr/bin/env python
"""Key counters of every kernel in an .ncu-rep (ncu -i ... --page raw --csv): duration, DRAM bytes, issue / LSU / tensor
utilisation, occupancy, stall reasons per issue.  Usage: python scripts/ncu_summary.py file.ncu-rep [more metrics...]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:100])
        for i, h in enumerate(hdr):
            if h in KEYS or any(e in h for e in extra) or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")) or "tensor" in h and "pct" in h:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if "issue_stalled" in h and v < 0.05:
                    continue
                print(f"  {h} [{units[i]}] = {r[i]}")


if __name__ == "__main__":
    main()
