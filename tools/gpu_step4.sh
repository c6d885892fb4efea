#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02h4}
timeout 900 python -m pytest tests/test_resnet.py tests/test_wider_eval.py -m gpu -q -p no:cacheprovider -s 2>&1 | grep -vE "^$" | tail -12
timeout 300 python tools/resnet_level.py 1408 1 2>&1 | tail -14 | tee gpurun_out/${T}_resnet_level1408.txt
