#!/bin/bash
# r02f: compute-sanitizer over the kernels written / rewritten in the second half of round 2 (conv1_tc halo staging,
# eltwise, general max pooling, first conv, strided 1x1 conv) + the new ResNet-through-caffe.Net test
mkdir -p gpurun_out
T=${TAG:-r02f}
timeout 600 python -m pytest tests/test_resnet.py -m gpu -q -p no:cacheprovider -k "caffe_net_surface" -s 2>&1 | tail -4 > gpurun_out/${T}_resnet_caffe_net.log; cat gpurun_out/${T}_resnet_caffe_net.log
echo "# compute-sanitizer --tool memcheck over the ResNet-op kernel tests and the conv1_1 tests (B200, final code)" > gpurun_out/${T}_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
   python -m pytest tests/test_resnet.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider \
   -k "eltwise_sum_kernel or maxpool_general or conv_first or strided_1x1 or conv1_tensor_core or conv1_c3" 2>&1 | tail -8 >> gpurun_out/${T}_sanitizer_memcheck.log
echo "memcheck rc=${PIPESTATUS[0]}" >> gpurun_out/${T}_sanitizer_memcheck.log
tail -5 gpurun_out/${T}_sanitizer_memcheck.log
echo "# compute-sanitizer --tool racecheck: shared-memory hazards of conv1_tc (halo staging) and conv_first (B200, final code)" > gpurun_out/${T}_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 \
   python -m pytest tests/test_resnet.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider \
   -k "conv_first or conv1_tensor_core" 2>&1 | tail -8 >> gpurun_out/${T}_sanitizer_racecheck.log
echo "racecheck rc=${PIPESTATUS[0]}" >> gpurun_out/${T}_sanitizer_racecheck.log
tail -5 gpurun_out/${T}_sanitizer_racecheck.log
