#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02h3}
timeout 900 python -m pytest tests/test_resnet.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
timeout 300 python tools/resnet_level.py 1408 1 2>&1 | tail -14 | tee gpurun_out/${T}_resnet_level1408.txt
timeout 600 python tools/level_conv_only.py 2048 5 2>/dev/null | tail -1
