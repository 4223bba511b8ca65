"""Summarise an .ncu-rep (one kernel launch, --set full) into a text file for profiles/.
usage: ncu_summarize.py <report.ncu-rep> <cell_steps_of_the_launch> <flop_per_cell_step> <out.txt>"""
import collections, csv, io, subprocess, sys

rep, cellsteps, flop_cs, out = sys.argv[1], float(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
val = {h: (v, u) for h, u, v in zip(hdr, units, r)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]
L = []
L.append(f"# ncu summary of {rep}  (ncu --set full --clock-control none --import-source on; one launch)")
for k in keys:
    if k in val:
        L.append(f"{k:75s} {val[k][0]} {val[k][1]}")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        v = float(val[h][0] or 0)
        if v >= 0.05:
            L.append(f"{h:75s} {v:.3f}")
ms = float(val["gpu__time_duration.sum"][0].replace(",", ""))
u = val["gpu__time_duration.sum"][1]
ms *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]
L.append(f"\nalgorithmic: {cellsteps:.4g} cell-steps x {flop_cs:g} FLOP = {cellsteps*flop_cs/1e12:.3f} TFLOP in {ms:.2f} ms (under ncu, cold) "
         f"= {cellsteps*flop_cs/ms/1e9:.3f} TFLOP/s")
dr = val.get("dram__bytes_read.sum"); dw = val.get("dram__bytes_write.sum")
L.append(f"traffic (dram read + write) per launch: {dr[0]} {dr[1]} + {dw[0]} {dw[1]}")

rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]; ix = {h: i for i, h in enumerate(h2)}
ops = collections.Counter(); samp = collections.Counter(); tot = stot = 0
lines = []
for k, rr in enumerate(rows[2:]):
    if len(rr) < len(h2): continue
    toks = rr[ix["Source"]].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0] if not op.startswith(("LDS", "STS", "LDL", "STL", "LDG", "STG", "MUFU")) else ".".join(op.split(".")[:2])
    n = float(rr[ix["Instructions Executed"]] or 0); s = float(rr[ix["# Samples"]] or 0)
    ops[op] += n; samp[op] += s; tot += n; stot += s
    lines.append((s, rr))
L.append(f"\ninstruction mix: {tot:.3e} warp-instructions = {tot*32/cellsteps:.1f} thread-instruction slots per cell-step")
for op, n in ops.most_common(22):
    L.append(f"  {op:14s} {n*32/cellsteps:8.2f} /cell-step  {100*n/tot:5.1f}% inst   {100*samp[op]/max(stot,1):5.1f}% stall samples")
stalls = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
L.append("\ntop stall locations (share of warp-stall samples, SASS, dominant reasons):")
lines.sort(key=lambda t: -t[0])
for s, rr in lines[:10]:
    st = sorted(((float(rr[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    L.append(f"  {100*s/max(stot,1):5.1f}%  {rr[ix['Source']][:64]:64s} {[(n, int(v)) for v, n in st]}")
open(out, "w").write("\n".join(L) + "\n")
print("\n".join(L[:40]))
