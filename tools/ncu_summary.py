#!/usr/bin/env python3
"""One-screen summary of an ncu report (raw page): duration, DRAM traffic, issue / pipe utilisation, occupancy, stalls.
    python tools/ncu_summary.py rep.ncu-rep [title]   -> text for profiles/"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
        ("launch__occupancy_limit_registers", "occupancy limit: registers (blocks)"), ("launch__occupancy_limit_shared_mem", "occupancy limit: shared mem (blocks)"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
        ("smsp__inst_executed_op_local_ld.sum", "local (spill) loads"), ("smsp__inst_executed_op_shared_ld.sum", "shared loads"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts")]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        vals = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("%s" % (sys.argv[2] if len(sys.argv) > 2 else rep))
        print("kernel: %s" % vals.get("Kernel Name", "?")[:200])
        for k, name in WANT:
            if k in vals and vals[k] != "":
                print("  %-42s %s %s" % (name, vals[k], u.get(k, "")))
        st = []
        for s in ("long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "not_selected", "branch_resolving", "no_instruction", "dispatch_stall",
                  "barrier", "mio_throttle", "lg_throttle", "tex_throttle", "drain", "imc_miss", "membar", "sleeping"):
            k = STALLS % s
            if k in vals and vals[k] not in ("", "0"):
                st.append((float(vals[k]), s))
        print("  stall cycles per issued instruction: " + ", ".join("%s %.2f" % (s, v) for v, s in sorted(st, reverse=True)[:8]))


if __name__ == "__main__":
    main()
