#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the numbers DESIGN.md / bench.py quote.

usage: ncu -i gpurun_out/full_chain.ncu-rep --page raw --csv | python tools/summarize_ncu.py > profiles/r02_ncu_full_chain_train.txt

One block per captured launch: `  <metric> <value> <unit>` lines (bench.py's ncu_traffic() parses the dram__bytes lines of
this format).  Metrics are selected by exact name or by pattern (ncu's names differ between chips and versions - the
round-1 version of this script asked for names this ncu does not emit and silently printed nothing for the tensor pipe);
pattern matches are printed only when non-zero.
"""
import csv
import re
import sys

EXACT = [
    "Kernel Name", "Grid Size", "Block Size",
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.max",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum",
    "sm__inst_executed.sum.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
PATTERNS = [
    r"^sm__pipe_tensor.*cycles_active.*\.avg\.pct_of_peak_sustained_(active|elapsed)$",
    r"^(TPC|SM_[A-C])\.TriageCompute\.sm(sp)?__pipe_tensor.*cycles_active.*\.avg$",
    r"^sm__ops_path_tensor_op_utc.*\.sum$",
    r"^sm__inst_executed_pipe_(tmem|uniform|tc|xu|fma|fmaheavy|fmalite|alu|lsu|uniform)\.(sum|avg\.pct_of_peak_sustained_active)$",
    r"^smsp__inst_executed_pipe_(tmem|xu|lsu|fma|alu|uniform)\.sum$",
    r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$",
    r"^smsp__pcsamp_warps_issue_stalled_(?!.*not_issued).*$",
    r"^smsp__pcsamp_sample_buffer.*$",
]


def nonzero(v):
    try:
        return float(v.replace(",", "")) != 0.0
    except ValueError:
        return v != ""


def main():
    rows = list(csv.reader(sys.stdin))
    if len(rows) < 3:
        print("no kernels in input")
        return
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    pats = [re.compile(p) for p in PATTERNS]
    extra = [h for h in hdr if any(p.search(h) for p in pats)]
    for n, r in enumerate(rows[2:]):
        print(f"== launch {n}")
        for w in EXACT:
            if w in idx and idx[w] < len(r) and r[idx[w]] != "":
                print(f"  {w:96s} {r[idx[w]]} {units[idx[w]]}")
        for w in extra:
            i = idx[w]
            if i < len(r) and nonzero(r[i]):
                print(f"  {w:96s} {r[i]} {units[i]}")


if __name__ == "__main__":
    main()
