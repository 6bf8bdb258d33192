"""Summarise a .ncu-rep (one kernel capture, --set full) into a small JSON + text record for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/NAME"""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
rec = {"report": rep}
for i, name in enumerate(hdr):
    if name in ("Kernel Name",) or name in KEYS:
        rec[name] = vals[i] + ((" " + units[i]) if units[i] else "")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = srows[1]; data = srows[2:]
ci = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[ci["# Samples"]]) for r in data) or 1
agg = {n[6:]: sum(int(r[ci[n]]) for r in data) for n in stalls}
rec["stall_samples_total"] = tot
rec["stall_breakdown_pct"] = {k: round(100.0 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v}
ops = {}
for r in data:
    op = r[ci["Source"]].split()[0] if r[ci["Source"]].split() else ""
    if op.startswith("@"):
        op = r[ci["Source"]].split()[1]
    ops[op] = ops.get(op, 0) + int(r[ci["Instructions Executed"]])
rec["warp_instructions_by_opcode_top"] = dict(sorted(ops.items(), key=lambda kv: -kv[1])[:25])
top = sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:15]
rec["hottest_instructions"] = [{"sass": r[ci["Source"]].strip(), "samples_pct": round(100.0 * int(r[ci["# Samples"]]) / tot, 1)} for r in top]
json.dump(rec, open(out + ".json", "w"), indent=1)
print(json.dumps({k: rec[k] for k in rec if k not in ("warp_instructions_by_opcode_top", "hottest_instructions")}, indent=1))
