#!/usr/bin/env python
"""Instruction-count summary of the built library's SASS (cuobjdump -sass), per kernel: the
mnemonics that prove what the kernels use -- UBLKCP (1-D TMA bulk copies), SYNCS (mbarrier
arrive / try_wait), ACQBULK / PREEXIT (programmatic dependent launch), FFMA, LDS / STS, LDG / STG,
SHFL, HMMA / UTCMMA (none: no tensor cores on this path).

    python tools/sass_summary.py > profiles/rN_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lsp-dsp-units_b200", "libb200conv.so")
WATCH = ["UBLKCP", "UTMALDG", "SYNCS", "ACQBULK", "PREEXIT", "FFMA", "FMUL", "FADD", "DADD", "LDS", "STS", "LDG", "STG",
         "LDGSTS", "SHFL", "BAR", "MEMBAR", "ATOM", "RED", "NANOSLEEP", "HMMA", "UTCMMA", "LDL", "STL"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kernels = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("b200conv::", "")
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + ".") or (w in ("ATOM", "RED") and op.startswith(w)):
                kernels[cur][w] += 1

print("SASS summary of %s (cubins: %s)" % (os.path.relpath(LIB, ROOT), ", ".join(arch)))
print("%-34s %7s " % ("kernel", "instrs") + " ".join("%7s" % w[:7] for w in WATCH))
tot = collections.Counter()
for k, c in kernels.items():
    print("%-34s %7d " % (k[:34], c["_total"]) + " ".join("%7d" % c[w] for w in WATCH))
    tot.update(c)
print("%-34s %7d " % ("ALL", tot["_total"]) + " ".join("%7d" % tot[w] for w in WATCH))
