# ncu evidence for the 4K10 inter clip (c3): launch list + --set full of one kernel family (run on the GPU box through gpurun)
# usage: bash tools/ncu_capture_c3.sh TAG KERNEL_REGEX [SKIP] [COUNT]
TAG=${1:-r2}; KREGEX=${2:-inter_pred}; SKIP=${3:-12}; COUNT=${4:-2}
set -x
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-per-config --workload c3_4k10_inter"
timeout 300 $B > gpurun_out/plain_${TAG}_c3.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c $COUNT -f -o gpurun_out/prof_${TAG}_$KREGEX $B > gpurun_out/ncu_${TAG}_c3_b.log 2>&1
ls -la gpurun_out/prof_${TAG}_$KREGEX.ncu-rep
