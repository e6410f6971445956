"""Aggregate the per-instruction stall samples of an ncu source page (SASS view) by CUDA source line.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all build/nuts_kernel.o ; nvdisasm -g -c nuts_kernel.sm_100a.cubin > disasm.txt
    python tools/ncu_lines.py src.csv disasm.txt <mangled-kernel-substring> [top]
"""
import collections
import csv
import re
import sys

src_csv, disasm, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
amap, cur, on = {}, None, False
for l in open(disasm):
    if l.startswith(".text."):
        on = kern in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m:
        amap.setdefault(int(m.group(1), 16), (cur, m.group(2).strip()))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
agg, aggi, aggt = collections.Counter(), collections.Counter(), collections.Counter()
stall = collections.defaultdict(collections.Counter)
keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot, base = 0, None
for r in rows[2:]:
    addr = int(r[ix["Address"]], 16)
    base = addr if base is None else base
    s = int(r[ix["# Samples"]] or 0)
    tot += s
    ln = amap.get(addr - base, (None, ""))[0]
    agg[ln] += s
    aggi[ln] += float(r[ix["Instructions Executed"]] or 0)
    aggt[ln] += float(r[ix["Thread Instructions Executed"]] or 0)
    for k in keys:
        stall[ln][k] += float(r[ix[k]] or 0)
byfile = collections.Counter()
for k, v in agg.items():
    byfile[k[0] if k else None] += v
print("samples", tot, {k: round(v / tot, 3) for k, v in byfile.most_common()})
for k, v in agg.most_common(top):
    st = ", ".join(f"{a[6:]}={int(b)}" for a, b in stall[k].most_common(3))
    print(f"{v:7d} {v / tot * 100:5.1f}%  {aggi[k] / 1e6:8.1f}M inst  {aggt[k] / max(aggi[k], 1):5.1f} thr  {k}  [{st}]")
