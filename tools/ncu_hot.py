"""Per-source-line instruction counts from an ncu report captured with --import-source on (needs -lineinfo):
   ncu -i REP --page source --print-source cuda,sass --csv [--launch-skip N --launch-count 1] > src.csv ; python tools/ncu_hot.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
lines = {}
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] == "" or not r[0].isdigit():      # SASS rows under a source line have an empty line number
        continue
    try:
        n = int(r[hdr.index("Instructions Executed")])
        smp = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    a = lines.setdefault(key, [0, 0, r[1].strip()])
    a[0] += n
    a[1] += smp
tot = sum(v[0] for v in lines.values())
tots = sum(v[1] for v in lines.values())
print(f"total warp instructions {tot}, samples {tots}")
for (f, l), (n, s, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n:>11} {100 * n / max(1, tot):5.1f}%  smp {100 * s / max(1, tots):5.1f}%  {f}:{l}: {src[:100]}")
