#!/usr/bin/env python
"""Registers / stack / spills / shared memory per kernel from the ptxas logs of the last build
(aquaticgymenv_b200/csrc/*.ptxas.log).  usage: python tools/ptxas_summary.py [substring ...]"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, "
                 r"(\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?")
want = sys.argv[1:]
for f in sorted(glob.glob(os.path.join(ROOT, "aquaticgymenv_b200", "csrc", "*.ptxas.log"))):
    ents = pat.findall(open(f).read())
    names = subprocess.run(["c++filt"] + [e[0] for e in ents], capture_output=True, text=True).stdout.splitlines()
    for (name, stack, ss, sl, regs, smem), d in zip(ents, names):
        d = d.replace("aqua::", "").replace("(StepArgs)", "").replace("void ", "")
        if want and not any(w in d for w in want):
            continue
        print("%-62s regs %3s stack %4s spill %s/%s smem %s" % (d[:62], regs, stack, ss, sl, smem))
