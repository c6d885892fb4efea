#!/bin/bash
# compute-sanitizer over the kernels written / rewritten in the second half of round 2: conv1_tc halo staging, the templated
# first-conv tensor-core kernel (7x7/2, 3x3/1), eltwise, general max pooling, SIMT first conv, strided 1x1 conv, the conv
# epilogue with the fused residual add (staged row loads)
mkdir -p gpurun_out
T=${TAG:-r02f}
SEL="eltwise_sum_kernel or maxpool_general or conv_first or strided_1x1 or conv1_tensor_core or conv1_c3 or conv7_tensor_core or fused_residual"
echo "# compute-sanitizer --tool memcheck over the kernel tests of tests/test_resnet.py and the conv1_1 tests (B200, final code)" > gpurun_out/${T}_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
   python -m pytest tests/test_resnet.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -k "$SEL" 2>&1 | tail -8 >> gpurun_out/${T}_sanitizer_memcheck.log
echo "memcheck rc=${PIPESTATUS[0]}" >> gpurun_out/${T}_sanitizer_memcheck.log
tail -4 gpurun_out/${T}_sanitizer_memcheck.log
echo "# compute-sanitizer --tool racecheck: shared-memory hazards of conv1_tc / conv_first_tc (patch staging, operand rows) and of the residual epilogue's staged row loads (B200, final code)" > gpurun_out/${T}_sanitizer_racecheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 \
   python -m pytest tests/test_resnet.py tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider \
   -k "conv_first or conv1_tensor_core or conv7_tensor_core or fused_residual" 2>&1 | tail -8 >> gpurun_out/${T}_sanitizer_racecheck.log
echo "racecheck rc=${PIPESTATUS[0]}" >> gpurun_out/${T}_sanitizer_racecheck.log
tail -4 gpurun_out/${T}_sanitizer_racecheck.log
timeout 600 python -m pytest tests/test_config_native.py tests/test_wider_eval.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
