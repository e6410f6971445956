"""Top stalled SASS instructions of an ncu source page, with the instructions that precede them.

    ncu -i prof.ncu-rep --page source --csv > src.csv
    python tools/ncu_insts.py src.csv [stall column, default stall_long_sb] [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
col = sys.argv[2] if len(sys.argv) > 2 else "stall_long_sb"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 15
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(float(r[ix["# Samples"]] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("samples", int(tot), {h[6:]: round(sum(float(r[ix[h]] or 0) for r in data) / tot, 3) for h in stall_cols
                            if sum(float(r[ix[h]] or 0) for r in data) / tot > 0.01})
order = sorted(range(len(data)), key=lambda i: -float(data[i][ix[col]] or 0))[:top]
for i in order:
    r = data[i]
    print(f"---- {r[ix['Source']][:72]:72s} {col[6:]} {float(r[ix[col]] or 0):8.0f}  samples {float(r[ix['# Samples']] or 0):8.0f}  "
          f"thr {r[ix['Avg. Threads Executed']]}")
    for j in range(max(0, i - 4), i):
        print("         ", data[j][ix["Source"]][:90])
