#!/usr/bin/env python
"""profiles/pipes.json + profiles/traffic.json from the summary of the step-kernel capture:
tools/ncu_pipes.py profiles/r2_step_final_summary.csv [steps in the profiled launch = 64] [envs = 524288]"""
import csv, json, sys

src = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
envs = int(sys.argv[3]) if len(sys.argv) > 3 else 524288
rows = list(csv.reader(open(src)))
hdr, units, val = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}


def f(name):
    return float(val[col[name]])


def to_bytes(name):
    u = units[col[name]].lower()
    return f(name) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


kernel = "k_env_step<2,1,1> (MODE_SAMPLED, TILE 1: shared-memory slot tile, SELFPLAY; %d steps in the profiled launch)" % steps
dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
pipes = {
    "kernel": kernel, "envs": envs,
    "alu_pipe_pct_of_peak": round(f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 2),
    "fma_pipe_pct_of_peak": round(f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), 2),
    "xu_pipe_pct_of_peak": round(f("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"), 2),
    "lsu_pipe_pct_of_peak": round(f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), 2),
    "issue_slots_busy_pct": round(f("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
    "active_lanes_per_instruction": round(f("smsp__thread_inst_executed_per_inst_executed.ratio"), 2),
    "warp_instructions_per_step": f("smsp__inst_executed.sum") / steps,
    "dram_pct_of_peak": round(f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), 2),
    "registers": int(f("launch__registers_per_thread")),
    "resident_blocks_per_sm": int(f("launch__occupancy_limit_registers")),
    "stall_no_instruction_per_issue": round(f("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"), 2),
    "stall_wait_per_issue": round(f("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"), 2),
    "stall_long_scoreboard_per_issue": round(f("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"), 2),
    "source": src + " (ncu --set full)",
}
traffic = {
    "kernel": kernel, "envs": envs, "steps_in_profiled_launch": steps,
    "dram_bytes_per_step": dram / steps,
    "note": "per env step: %.0f B (one repetition-table entry read per ply, fetched by L2 as 64-128 B; state and slots stay on chip inside a multi-step launch)" % (dram / steps / envs),
    "source": src + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one %d-step launch over all %s envs)" % (steps, format(envs, ",")),
}
json.dump(pipes, open("profiles/pipes.json", "w"), indent=1)
json.dump(traffic, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(pipes, indent=1)); print(json.dumps(traffic, indent=1))
