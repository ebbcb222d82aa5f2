#!/usr/bin/env python
"""SASS opcode histogram of the product library (cuobjdump -sass), per kernel, restricted to the mnemonics that
prove what the code is built from: tcgen05 (UTCHMMA / UTCBAR / LDTM / STTM), bulk-async copies (UBLKCP), mbarrier
(SYNCS), cp.async (LDGSTS), vector atomics (RED / ATOMG), FFMA / HFMA2 counts.

    python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mvsnerf_b200", "libmvsnerf_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEEP = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "REDG", "ATOMG", "FFMA", "FFMA2",
        "HFMA2", "HMNMX2", "F2FP", "MUFU", "LDG", "STG", "LDS", "STS", "BAR", "ELECT", "UCGABAR", "DFMA")
kern, hist, order = None, collections.OrderedDict(), []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").split("(")[0]
        hist.setdefault(kern, collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Za-z0-9_]+)*)", line)
    if m and kern:
        op, mods = m.group(1), m.group(2)
        hist[kern][op] += 1
        hist[kern]["*total"] += 1
        if op in ("UTCHMMA", "UTCBAR", "UBLKCP", "REDG", "ATOMG", "LDTM", "STTM") and mods:
            hist[kern][op + mods] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} : opcode counts per kernel (static instruction counts)")
for k, c in hist.items():
    if c["*total"] < 40:
        continue
    sel = {o: n for o, n in c.items() if o.split(".")[0] in KEEP}
    print(f"\n{k}   [{c['*total']} instructions]")
    print("   " + "  ".join(f"{o}={n}" for o, n in sorted(sel.items())))
tot = collections.Counter()
for c in hist.values():
    tot.update(c)
print("\n# library totals: " + "  ".join(f"{o}={tot[o]}" for o in sorted(tot) if o.split(".")[0] in
                                         ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "REDG")))
