#!/bin/bash
# Multi-GPU evidence on one box: the bench line under torchrun (weak scaling + the plugin leg per rank) with BASELINE
# configs[4] (3226 WIDER-val-shaped images, reference partition), then the area-sorted partition.   usage: gpu_multi.sh N
N=${1:-2}
T=r02
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${T}_multi${N}_gpus.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 1500 $RUN bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --wider-shaped 3226 \
    > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
tail -c 1800 gpurun_out/${T}_bench_${N}gpu.json; tail -3 gpurun_out/${T}_bench_${N}gpu.err
timeout 900 $RUN tools/wider_shaped_run.py --images 3226 --partition area_rr \
    > gpurun_out/${T}_wider_area_rr_${N}gpu.json 2> gpurun_out/${T}_wider_${N}gpu.err
tail -c 900 gpurun_out/${T}_wider_area_rr_${N}gpu.json; tail -3 gpurun_out/${T}_wider_${N}gpu.err
