"""Dev tool: aggregate `ncu --page source --csv --print-source cuda,sass` output per source line.
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > src.csv; python tools/ncu_lines.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
agg = {}
total = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_s = hdr.index("Warp Stall Sampling (All Samples)")
        i_i = hdr.index("Instructions Executed")
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0].isdigit():  # a source line row (aggregated over its SASS)
        try:
            s, n = int(r[i_s]), int(r[i_i])
        except ValueError:
            continue
        key = (cur_file, int(r[0]), r[1].strip()[:110])
        a = agg.setdefault(key, [0, 0])
        a[0] += s
        a[1] += n
        total += s
print(f"total samples {total}")
for (f, ln, src), (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100.0 * s / max(total, 1):5.1f}%  {s:7d}  inst {n:9d}  {f}:{ln}  {src}")
