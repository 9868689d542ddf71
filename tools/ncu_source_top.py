"""Top stall sites from an `ncu --page source --csv` dump. usage: ncu_source_top.py file.csv [kernel_index] [N]"""
import csv
import sys

csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
# split into kernels
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
s, e = starts[kidx], starts[kidx + 1]
print(rows[s][1][:120])
hdr = rows[s + 1]
ci = {h: i for i, h in enumerate(hdr)}
body = rows[s + 2:e]
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
ranked = sorted(enumerate(body), key=lambda t: -int(t[1][ci["# Samples"]] or 0))[:N]
for idx, r in sorted(ranked):
    print(f"{idx:5d} {int(r[ci['# Samples']]):7d} {100*int(r[ci['# Samples']])/tot:5.1f}%  exec={r[ci['Instructions Executed']]:>8}  {r[ci['Source']].strip()[:110]}")
