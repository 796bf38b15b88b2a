"""Opcode histogram per kernel of libbdlm.so (cuobjdump -sass), committed under profiles/.

    python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt

What to look for (B200_PROFILING.md): LDGSTS = cp.async staging, UTMALDG / UBLKCP = TMA,
UTC*MMA / LDTM = tcgen05, HMMA / DMMA = legacy tensor path.  This path is fp64 recursions at
0.6 flop/B (config 2) or small dense fp64 chains (configs 3, 4): no tensor-core instruction is
expected; the FP64 pipe (DADD / DMUL / DFMA) and the memory pipes are what matter."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "bayesian_dlms_b200", "libbdlm.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::|bdlm::", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
FLAG = ("LDGSTS", "UTMALDG", "UTMASTG", "UBLKCP", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "HMMA", "DMMA")
print("# cuobjdump -sass of", os.path.relpath(so, ROOT), "(sm_100a): static instruction counts per kernel")
print("# columns: total | FP64 (DADD+DMUL+DFMA+DSETP+MUFU) | LDG/STG | LDS/STS | SHFL | LDL/STL | flagged opcodes")
for k, h in hist.items():
    tot = sum(h.values())
    if tot == 0:
        continue
    f64 = sum(h[o] for o in ("DADD", "DMUL", "DFMA", "DSETP", "MUFU"))
    g = h["LDG"] + h["STG"] + h["LD"] + h["ST"]
    sh = h["LDS"] + h["STS"]
    lo = h["LDL"] + h["STL"]
    flags = " ".join(f"{o}={h[o]}" for o in FLAG if h[o])
    print(f"{tot:7d} | {f64:6d} | {g:5d} | {sh:5d} | {h['SHFL']:5d} | {lo:5d} | {flags or '-':24s} | {k[:110]}")
