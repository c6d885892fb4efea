#!/bin/bash
# ncu --set full captures of the kernels added for residual backbones (one launch each, 1408x1408 level of the ResNet detector)
mkdir -p gpurun_out
T=r02f
N="ncu --set full --clock-control none --import-source on -f"
timeout 600 $N -k regex:conv_first_tc_kernel -s 1 -c 1 -o gpurun_out/${T}_resnet_conv_first_tc python tools/resnet_level.py 1408 1 > /dev/null 2>&1
timeout 600 $N -k regex:maxpool_h2_kernel -s 1 -c 1 -o gpurun_out/${T}_resnet_maxpool3x3 python tools/resnet_level.py 1408 1 > /dev/null 2>&1
timeout 600 $N -k regex:eltwise_sum_kernel -s 8 -c 1 -o gpurun_out/${T}_resnet_eltwise python tools/resnet_level.py 1408 1 nofuse > /dev/null 2>&1
timeout 600 $N --kernel-name-base demangled -k "regex:conv_stream_kernel<.*true" -s 8 -c 1 -o gpurun_out/${T}_resnet_conv_res python tools/resnet_level.py 1408 1 > /dev/null 2>&1
ls -la gpurun_out/${T}_resnet_*.ncu-rep
