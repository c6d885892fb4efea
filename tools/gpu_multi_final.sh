#!/bin/bash
# final-code multi-GPU sanity: the bench line under torchrun + a short WIDER-shaped run.   usage: gpu_multi_final.sh N [images]
N=${1:-2}
IM=${2:-512}
T=r02f
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 1200 $RUN bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --wider-shaped $IM \
    > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_${N}gpu.json').read().strip().splitlines()[-1])
w=d.get('wider_shaped') or {}
print('N=%d value %.1f ms %.2f frac %.4f e2e %.1f batched %.1f | wider %s img/s imbalance %s busy %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e_batched']['value'], w.get('value'), w.get('imbalance'), w.get('rank_busy_seconds')))
PY
tail -3 gpurun_out/${T}_bench_${N}gpu.err
timeout 600 $RUN bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -c 400
