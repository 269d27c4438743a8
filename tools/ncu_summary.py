"""Summarise an .ncu-rep (raw page) into a small JSON for profiles/: per kernel duration, registers, FP64 pipe %, DRAM bytes,
executed FP64 instruction counts → flop per element when --nele is given."""
import csv, json, subprocess, sys
rep = sys.argv[1]; nele = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
keys = {"duration_ms": "gpu__time_duration.sum", "registers": "launch__registers_per_thread", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum", "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "inst_executed": "smsp__inst_executed.sum", "dfma": "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "dmul": "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "dadd": "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "local_ld": "smsp__sass_inst_executed_op_local_ld.sum", "local_st": "smsp__sass_inst_executed_op_local_st.sum",
        "stall_no_inst": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "stall_long_sb": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_math": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"}
units = rows[1]
res = []
for r in rows[2:]:
    d = {"kernel": r[idx["Kernel Name"]][:90]}
    for k, m in keys.items():
        if m in idx:
            try:
                d[k] = float(r[idx[m]]); d[k + "_unit"] = units[idx[m]]
            except ValueError:
                pass
    cyc = None
    for m in ("sm__cycles_elapsed.max", "smsp__cycles_elapsed.max"):
        if m in idx:
            try: cyc = float(r[idx[m]]); break
            except ValueError: pass
    for k in ("dfma", "dmul", "dadd"):
        m = "smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % k
        if k not in d and m in idx and cyc:
            try: d[k] = float(r[idx[m]]) * cyc; d[k + "_unit"] = "inst (per_cycle_elapsed x sm__cycles_elapsed.max)"
            except ValueError: pass
    if nele and "dfma" in d and "beam_kernel" in d["kernel"]:
        d["flop_per_element"] = (2 * d["dfma"] + d.get("dmul", 0) + d.get("dadd", 0)) / nele
        d["fp64_inst_per_element"] = (d["dfma"] + d.get("dmul", 0) + d.get("dadd", 0)) / nele
    res.append(d)
print(json.dumps(res, indent=1))
