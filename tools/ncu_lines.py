"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
    ncu -i x.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_lines.py src.csv [N]
Prints the N hottest lines by warp-stall samples with executed warp instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = ""
agg = {}
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or not r[0] or not r[0].isdigit():
        continue
    try:
        s, e = int(r[i_s]), int(r[i_e])
    except ValueError:
        continue
    k = (cur_file, int(r[0]), r[1].strip()[:110])
    a = agg.setdefault(k, [0, 0])
    a[0] += s
    a[1] += e
tot_s = sum(a[0] for a in agg.values())
tot_e = sum(a[1] for a in agg.values())
print(f"total samples {tot_s}  warp instructions {tot_e}")
for (f, ln, src), (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * s / tot_s:5.1f}% {s:8d} {100 * e / tot_e:5.1f}%i {f}:{ln}  {src}")
