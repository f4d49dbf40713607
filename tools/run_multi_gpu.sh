# Multi-GPU checks on an N-GPU box (gpurun --gpus N): in-process pool test + the torchrun bench line the driver uses.
# usage: bash tools/run_multi_gpu.sh TAG N
TAG=$1; N=$2
nvidia-smi -L > gpurun_out/multi_${TAG}_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_batch.py -m gpu -x -q > gpurun_out/multi_${TAG}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/multi_${TAG}_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/multi_${TAG}_bench_n$N.json 2> gpurun_out/multi_${TAG}_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/multi_${TAG}_ref_n$N.json 2> gpurun_out/multi_${TAG}_ref_n$N.err
tail -2 gpurun_out/multi_${TAG}_pytest.log; tail -c 600 gpurun_out/multi_${TAG}_bench_n$N.json
