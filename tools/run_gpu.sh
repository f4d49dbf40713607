# usage: bash tools/run_gpu.sh TAG [workloads...]   -- GPU parity tests + bench lines into gpurun_out/
TAG=$1; shift
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
for w in "$@"; do
  AV1R_PROFILE=1 timeout 600 python bench.py --steps 5 --warmup 3 --workload $w > gpurun_out/bench_${TAG}_$w.json 2> gpurun_out/bench_${TAG}_$w.err
  tail -2 gpurun_out/bench_${TAG}_$w.err
done
