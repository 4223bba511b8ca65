"""Aggregate an `ncu --page source --csv` export by SASS opcode: executed warp-instructions and stall samples.
usage: ncu_opmix.py src.csv <cell_steps>"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
cellsteps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samp = collections.Counter(); tot = 0; stot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0] if not op.startswith(("LDS", "STS", "LDL", "STL", "LDG", "STG", "MUFU")) else ".".join(op.split(".")[:2])
    n = float(r[ix["Instructions Executed"]] or 0); s = float(r[ix["# Samples"]] or 0)
    ops[op] += n; samp[op] += s; tot += n; stot += s
print(f"total warp-instr {tot:.3e}  = {tot*32/cellsteps:.1f} thread-instr per cell-step; SASS lines {len(rows)-2}")
for op, n in ops.most_common(40):
    print(f"{op:14s} {n*32/cellsteps:8.2f} /cell-step  {100*n/tot:5.1f}% inst   {100*samp[op]/max(stot,1):5.1f}% samples")
