#!/bin/bash
# conv1_tc after the halo-staging rewrite: parity tests, stand-alone timing, the level table
mkdir -p gpurun_out
T=${TAG:-r02h}
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "conv1 or roundtrip" 2>&1 | tail -5 > gpurun_out/${T}_c1_pytest.log
cat gpurun_out/${T}_c1_pytest.log
for f in 1 0; do timeout 300 python tools/time_conv1.py 2048 1 $f; timeout 300 python tools/time_conv1.py 1408 16 $f; done 2>&1 | tee gpurun_out/${T}_c1_time.txt
timeout 600 python tools/level_conv_only.py 2048 5 2>/dev/null | tail -3
