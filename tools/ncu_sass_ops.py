"""Dev tool: executed warp instructions and stall samples per SASS opcode from `ncu --page source --csv --print-source cuda,sass`.
    python tools/ncu_sass_ops.py src.csv [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ops = collections.defaultdict(lambda: [0, 0])
seen = set()
for r in rows:
    if len(r) < 8 or r[0] != '' or not r[2].startswith('0x'):
        continue
    if r[2] in seen:
        continue
    seen.add(r[2])
    toks = r[3].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0] if not op.startswith(('UTC', 'LDTM', 'STTM', 'SYNCS', 'UBLK')) else op
    try:
        ops[op][0] += int(r[7]); ops[op][1] += int(r[4])
    except ValueError:
        pass
ti = sum(v[0] for v in ops.values()); ts = sum(v[1] for v in ops.values())
print(f"instructions {ti}  samples {ts}")
for k, (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100.0 * n / ti:5.1f}%  {n:11d}  stall {100.0 * s / max(ts, 1):5.1f}%  {k}")
