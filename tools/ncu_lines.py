"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_lines.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
hdr = None
out = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
        continue
    if hdr and r and r[0] not in ("", "Function Name") and r[0].isdigit():
        try:
            out.append((cur, int(r[0]), r[1].strip()[:90], int(r[iS] or 0), int(r[iI] or 0)))
        except ValueError:
            pass
tot_s = sum(o[3] for o in out) or 1
tot_i = sum(o[4] for o in out) or 1
print(f"total samples {tot_s}, total warp instructions {tot_i}")
print("--- by samples")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{100*o[3]/tot_s:5.1f}% smp {100*o[4]/tot_i:5.1f}% inst  {o[0]}:{o[1]}  {o[2]}")
