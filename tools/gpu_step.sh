#!/bin/bash
# quick check of a conv-kernel change: kernel + ResNet parity tests, the 2048 level table, the ResNet level table
mkdir -p gpurun_out
T=${TAG:-r02l}
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_resnet.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
timeout 600 python tools/level_conv_only.py 2048 5 2>/dev/null > gpurun_out/${T}_level2048_conv_only.txt; grep -A30 "hf8" gpurun_out/${T}_level2048_conv_only.txt | grep -E "conv1_2|conv2_1|conv2_2|conv4_2|ALL"
timeout 600 python tools/resnet_level.py 1408 1 2>&1 | tail -16 | tee gpurun_out/${T}_resnet_level1408.txt
