#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02h5}
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_resnet.py -m gpu -q -p no:cacheprovider -s -k "conv1 or conv7 or resnet_net_forward" 2>&1 | grep -vE "^$|per-blob" | tail -12
for f in 1 0; do timeout 300 python tools/time_conv1.py 2048 1 $f; done 2>&1 | tee gpurun_out/${T}_conv1_time.txt
timeout 300 python tools/time_conv1.py 1408 16 1 2>&1 | tee -a gpurun_out/${T}_conv1_time.txt
for impl in single pair single pair; do
  SHF_CONV1_IMPL=$impl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('conv1 impl $impl: value %.2f ms %.2f frac %.4f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz']))" | tee -a gpurun_out/${T}_conv1_time.txt
done
