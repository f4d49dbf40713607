# ncu launch list (per-kernel gpu__time_duration, cold-cache + serialised: compare SHARES) for one workload
# usage: bash tools/ncu_launches.sh TAG WORKLOAD [SKIP] [COUNT]
TAG=$1; W=$2; SKIP=${3:-2000}; COUNT=${4:-800}
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-config --workload $W"
timeout 300 $B > gpurun_out/plain_${TAG}_$W.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $COUNT --csv --log-file gpurun_out/launches_${TAG}_$W.csv $B > gpurun_out/ncu_${TAG}_$W.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_${TAG}_$W.csv")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) != len(h): continue
    name = r[kn].split("(")[0].replace("void av1r::", "")
    a = agg[name]; a[0] += 1; a[1] += float(r[mv].replace(",", "")) / 1000.0
tot = sum(v[1] for v in agg.values())
with open("gpurun_out/launch_summary_${TAG}_$W.txt", "w") as f:
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:60s} {n:5d} launches {us / 1000:9.3f} ms  {100 * us / tot:5.1f}%  avg {us / n:8.1f} us\n")
PY
cat gpurun_out/launch_summary_${TAG}_$W.txt
