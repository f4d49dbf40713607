# What the driver runs at round end, N = 1: reference arm then our arm, default workload; wall time of each recorded.
TAG=$1
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/drv_${TAG}_ref.json 2> gpurun_out/drv_${TAG}_ref.err ) 2> gpurun_out/drv_${TAG}_ref.time
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/drv_${TAG}_n1.json 2> gpurun_out/drv_${TAG}_n1.err ) 2> gpurun_out/drv_${TAG}_n1.time
tail -3 gpurun_out/drv_${TAG}_n1.err; cat gpurun_out/drv_${TAG}_ref.time gpurun_out/drv_${TAG}_n1.time
