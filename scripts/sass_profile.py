#!/usr/bin/env python3
"""Static profile of a kernel's SASS (no GPU needed): basic blocks with instruction mix and the scheduler's stall
counts (the fixed-latency issue cycles ptxas encoded in the control bits), grouped by source line.

    python scripts/sass_profile.py build/classic_uniform.o 'classic_uniform_kernelILi13ELi8ELi16ELi168ELb1ELb0ELb1'
        [--blocks] [--lines] [--min-instr N]

For every basic block: address range, instruction count, FP64-pipe instructions (DFMA/DMUL/DADD/DSETP), shared /
local memory instructions, the sum of stall counts (= cycles one warp needs to issue the block when nothing else
delays it) and the source lines it comes from.  A block whose stall sum is far above 2 x (its FP64 instructions)
is latency-bound for a single warp: the other resident warps have to fill that gap.

Control bits (sm_70 and later, upper 64-bit word of the 128-bit instruction): stall = bits 41-44, yield = 45,
write barrier = 46-48, read barrier = 49-51, wait mask = 52-57.
"""
from __future__ import annotations

import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
INSTR = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;\s*/\* (0x[0-9a-f]{16}) \*/")
HI = re.compile(r"^\s*/\* (0x[0-9a-f]{16}) \*/")
LINE = re.compile(r'//## File "(.*?)", line (\d+)')
LABEL = re.compile(r"^(\.L_x_\d+):")


def disasm(obj: str) -> str:
    with tempfile.TemporaryDirectory() as tmp:
        if obj.endswith(".cubin"):
            cubin = obj
        else:
            subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
            cubins = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
            cubin = os.path.join(tmp, cubins[0])
        return subprocess.run(["nvdisasm", "-c", "-g", "-hex", cubin], check=True, capture_output=True, text=True).stdout


def parse(text: str, pattern: str):
    """-> list of dict(addr, op, text, line, stall, yield_, wbar, rbar, wait, label)"""
    out, on, cur_line, pending_label = [], False, 0, None
    rx = re.compile(pattern)
    for ln in text.splitlines():
        if ln.startswith("\t.section") or ln.startswith(".section"):
            on = ".text." in ln and rx.search(ln) is not None
            continue
        if not on:
            continue
        m = LABEL.match(ln)
        if m:
            pending_label = m.group(1)
            continue
        m = LINE.search(ln)
        if m:
            cur_line = int(m.group(2)) if m.group(1).endswith(".cu") else -int(m.group(2))
            continue
        m = INSTR.match(ln)
        if m:
            body = m.group(2)
            toks = body.split()
            pred = toks[0] if toks[0].startswith("@") else ""
            op = toks[1] if pred else toks[0]
            out.append(dict(addr=int(m.group(1), 16), op=op, text=body, line=cur_line, label=pending_label, pred=pred))
            pending_label = None
            continue
        m = HI.match(ln)
        if m and out and "stall" not in out[-1]:
            hi = int(m.group(1), 16)
            out[-1].update(stall=(hi >> 41) & 0xF, yield_=(hi >> 45) & 1, wbar=(hi >> 46) & 7, rbar=(hi >> 49) & 7,
                           wait=(hi >> 52) & 0x3F)
    return out


def blocks(ins):
    """split at labels and after control transfers"""
    bl, cur = [], []
    for i in ins:
        if i["label"] and cur:
            bl.append(cur); cur = []
        cur.append(i)
        if i["op"].split(".")[0] in ("BRA", "EXIT", "RET", "BRX", "JMP", "CALL", "BSYNC", "WARPSYNC", "BAR"):
            bl.append(cur); cur = []
    if cur:
        bl.append(cur)
    return bl


def kind(op: str) -> str:
    b = op.split(".")[0]
    if b in FP64:
        return "fp64"
    if b in ("LDS", "STS", "LDSM"):
        return "smem"
    if b in ("LDL", "STL"):
        return "local"
    if b in ("LDG", "STG", "LD", "ST", "ATOMG", "RED", "LDC", "LDCU"):
        return "gmem"
    if b in ("SHFL",):
        return "shfl"
    if b in ("MUFU",):
        return "mufu"
    if b in ("BRA", "BSSY", "BSYNC", "BAR", "WARPSYNC", "EXIT"):
        return "ctrl"
    return "alu"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("pattern")
    ap.add_argument("--blocks", action="store_true", help="print every basic block")
    ap.add_argument("--lines", action="store_true", help="aggregate by source line")
    ap.add_argument("--min-instr", type=int, default=8)
    ap.add_argument("--dump", type=str, default="", help="print the instructions of the block that starts at this hex address")
    a = ap.parse_args()
    ins = parse(disasm(a.obj), a.pattern)
    if not ins:
        sys.exit("no instructions matched")
    tot = collections.Counter(kind(i["op"]) for i in ins)
    print(f"# {len(ins)} instructions; mix {dict(tot)}; stall sum {sum(i.get('stall', 0) for i in ins)}")
    bl = blocks(ins)
    if a.blocks:
        print("# addr_lo-addr_hi  n  fp64 smem local alu  stall  stall/fp64x2  lines  | last")
        for b in bl:
            if len(b) < a.min_instr:
                continue
            c = collections.Counter(kind(i["op"]) for i in b)
            st = sum(i.get("stall", 0) for i in b)
            lines = sorted({i["line"] for i in b if i["line"] > 0})
            lr = f"{lines[0]}-{lines[-1]}" if lines else "-"
            f2 = 2 * c["fp64"]
            print(f"{b[0]['addr']:06x}-{b[-1]['addr']:06x} {len(b):5d} {c['fp64']:5d} {c['smem']:4d} {c['local']:4d} {c['alu']:5d} "
                  f"{st:6d}  {st / f2 if f2 else 0:5.2f}  {lr:>9s} | {b[-1]['text'][:48]}")
    if a.lines:
        agg = collections.defaultdict(lambda: collections.Counter())
        for i in ins:
            k = agg[i["line"]]
            k["n"] += 1; k[kind(i["op"])] += 1; k["stall"] += i.get("stall", 0)
        print("# line  n  fp64 smem local alu stall")
        for line in sorted(agg):
            k = agg[line]
            print(f"{line:6d} {k['n']:5d} {k['fp64']:5d} {k['smem']:4d} {k['local']:4d} {k['alu']:5d} {k['stall']:6d}")
    if a.dump:
        start = int(a.dump, 16)
        for b in bl:
            if b[0]["addr"] == start:
                for i in b:
                    print(f"{i['addr']:06x} L{i['line']:<5d} st={i.get('stall', 0):2d} y={i.get('yield_', 0)} w={i.get('wbar', 7)} "
                          f"r={i.get('rbar', 7)} wt={i.get('wait', 0):02x}  {i['text']}")


if __name__ == "__main__":
    main()
