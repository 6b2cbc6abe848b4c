#!/usr/bin/env python
"""Selected raw metrics per kernel launch of an .ncu-rep:  python tools/ncu_metrics.py REPORT.ncu-rep [extra_metric ...]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg"]


def main(rep, extra):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print("kernels:", [r[ix["Kernel Name"]].replace("void ocgcn::", "")[:44] for r in data])
    for w in WANT + list(extra):
        if w in ix:
            print("%-70s %-10s %s" % (w, units[ix[w]], [r[ix[w]] for r in data]))
        else:
            print("missing:", w)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
