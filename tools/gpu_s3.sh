#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet.py -m gpu -q -p no:cacheprovider -s -k "conv3x3_stride2 or stride_on_the_3x3 or strided_1x1 or fused_residual" 2>&1 | grep -vE "^$" | tail -15 | tee gpurun_out/r02f_s3_tests.txt
