"""Opcode histogram per kernel of the built library (evidence that the hot loops are DFMA / DMMA code for sm_100a).

    python tools/sass_histogram.py [path/to/lib.so] > profiles/sass_opcodes.txt

FP64 has no tcgen05 form: the Blackwell tensor path for doubles is DMMA (mma.sync.m8n8k4.f64), which is what the
group kernels use; UTMALDG / UTCMMA / LDTM (TMA / tcgen05 / TMEM) are therefore expected to be absent.
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
so = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "smc-nuts_b200/smcnuts/_lib/libsmcnuts_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip() or s  # noqa: E731
kern, hist, order = None, {}, []
arch = set(re.findall(r"arch = (sm_\w+)", out))
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        order.append(kern)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        op = m.group(1)
        if op == "DMMA":
            op += m.group(2)
        hist[kern][op] += 1
print(f"library: {Path(so).name}   arch: {sorted(arch)}")
keys = ("DFMA", "DMUL", "DADD", "DMMA", "MUFU", "LDG", "STG", "LDS", "STS", "LDGSTS", "LDL", "STL", "SHFL", "IMAD", "BRA", "UTMALDG", "UTCMMA", "LDTM")
tot = collections.Counter()
for k in order:
    h = hist[k]
    n = sum(h.values())
    grouped = collections.Counter()
    for op, c in h.items():
        grouped[op.split(".")[0]] += c
        tot[op.split(".")[0]] += c
    name = demangle(k)
    print(f"\n{name}\n  instructions {n}: " + ", ".join(f"{kk} {grouped[kk]}" for kk in keys if grouped[kk]))
    dm = {op: c for op, c in h.items() if op.startswith("DMMA")}
    if dm:
        print("  " + ", ".join(f"{op} {c}" for op, c in dm.items()))
print("\nwhole library: " + ", ".join(f"{kk} {tot[kk]}" for kk in keys))
