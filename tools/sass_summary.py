#!/usr/bin/env python
"""SASS evidence for the shipped libaquagym.so (read here, no GPU): per kernel the opcode histogram
and the first occurrences of the Blackwell-specific instructions (packed FP32: FFMA2 / FADD2 / FMUL2;
bulk asynchronous copies: UBLKCP; mbarrier: SYNCS; cp.async: LDGSTS).
usage: python tools/sass_summary.py > profiles/r03_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "aquaticgymenv_b200", "libaquagym.so")
KERNELS = [
    ("step_many_kernel<float, 2 envs per thread, I32 tape, small table, ring drawn at the start> (headline, C2)", "_ZN4aqua16step_many_kernelIfLi2ELi0ELb1ELi1ELb0ELb0EEEvNS_8StepArgsE"),
    ("step_many_kernel<float, 2, F32X2 tape, small table, start> (C3)", "_ZN4aqua16step_many_kernelIfLi2ELi3ELb1ELi1ELb0ELb0EEEvNS_8StepArgsE"),
    ("step_many_kernel<float, 2, I32 tape, culling grid, ring drawn in the loop> (C4)", "_ZN4aqua16step_many_kernelIfLi2ELi0ELb1ELi2ELb1ELb0EEEvNS_8StepArgsE"),
    ("step_many_kernel<float, 1, I32, small table, in-loop, in-kernel Q-network> (aqua_rollout_policy)", "_ZN4aqua16step_many_kernelIfLi1ELi0ELb0ELi1ELb1ELb1EEEvNS_8StepArgsE"),
    ("step_stream_kernel<float, 4, I32, waves, bulk staging> (one step per launch, 2^22+ envs)", "_ZN4aqua18step_stream_kernelIfLi4ELi0ELb1ELb1EEEvNS_8StepArgsE"),
    ("step_stream_kernel<float, 4, I32, waves, cp.async staging>", "_ZN4aqua18step_stream_kernelIfLi4ELi0ELb1ELb0EEEvNS_8StepArgsE"),
    ("step_narrow_kernel<float, I32, waves>", "_ZN4aqua18step_narrow_kernelIfLi0ELb1EEEvNS_8StepArgsE"),
]
SPECIAL = ("FFMA2", "FADD2", "FMUL2", "UBLKCP", "SYNCS", "LDGSTS", "UTMALDG", "REDUX", "IDP", "MOV", "LDS", "LDCU")
print("SASS summary of %s (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -fmad=false)\n" % os.path.relpath(LIB, ROOT))
for title, sym in KERNELS:
    out = subprocess.run(["cuobjdump", "-sass", "-fun", sym, LIB], capture_output=True, text=True).stdout
    lines = [re.sub(r"/\*[0-9a-f]{16}\*/\s*$", "", l).rstrip() for l in out.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]
    ops = collections.Counter()
    for l in lines:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            ops[m.group(1)] += 1
    print("== %s\n   %s\n   %d instructions" % (title, sym, len(lines)))
    print("   Blackwell / async instructions: " + ", ".join("%s x%d" % (k, ops[k]) for k in SPECIAL if ops[k]))
    print("   top opcodes: " + ", ".join("%s %d" % kv for kv in ops.most_common(24)))
    shown = collections.Counter()
    for l in lines:
        for k in ("FFMA2", "FADD2", "FMUL2", "UBLKCP", "SYNCS"):
            if re.search(r"\b%s\b" % k, l) and shown[k] < 2:
                shown[k] += 1
                print("     " + l.strip())
    print()
