#!/usr/bin/env python
"""Condense an `ncu --set full` report into the per-kernel table kept under profiles/.

usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_x_summary.csv
Reads the report with `ncu -i … --page raw --csv` (works without a GPU) and keeps the columns the roofline
in bench.py / DESIGN.md is argued from: duration, DRAM bytes (-> roofline.traffic), L2 bytes / hit rate,
issue-slot use, achieved occupancy, registers, tensor-pipe activity, atomic (RED) traffic at L2.
"""
import csv
import subprocess
import sys

COLS = [
    ("Kernel Name", "kernel"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sectors_op_red.sum", "l2_red_sectors"),
    ("l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum", "red_sectors"),      # lane-reductions leaving the SM (hash scatter)
    ("smsp__inst_executed_op_global_red.sum", "red_insts"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sectors"),                # global-load sectors through L1 (hash gathers)
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld_requests"),
    ("lts__t_sectors_op_atom.sum", "l2_atom_sectors"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_insts"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    keep = [(hdr.index(m), short, units[hdr.index(m)]) for m, short in COLS if m in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([s + (f" [{u}]" if u else "") for _, s, u in keep])
        for r in body:
            w.writerow([r[i][:110] for i, _, _ in keep])
    print(f"{len(body)} kernels -> {out}")


if __name__ == "__main__":
    main()
