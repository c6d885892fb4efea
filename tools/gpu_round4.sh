#!/bin/bash
# r02d: ncu of the fused-pool conv1_2 after the epilogue rewrite, compute-sanitizer pass over the kernel tests
T=r02d
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_stream_kernel<.int.64" -s 3 -c 1 -f -o gpurun_out/${T}_conv1_2_pool_fmt1 \
   python tools/level_conv_only.py 2048 1 > gpurun_out/${T}_ncu1.log 2>&1
tail -2 gpurun_out/${T}_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_stream_kernel<.int.128" -s 39 -c 1 -f -o gpurun_out/${T}_conv2_2_pool_fmt1 \
   python tools/level_conv_only.py 2048 1 > gpurun_out/${T}_ncu2.log 2>&1
tail -2 gpurun_out/${T}_ncu2.log
# compute-sanitizer (memcheck) over the kernel-level tests: every CUDA kernel of the library with small shapes
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
   python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -x \
   -k "not ragged_batch and not hf8_matches_operand_model" > gpurun_out/${T}_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/${T}_sanitizer_memcheck.log
tail -6 gpurun_out/${T}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 \
   python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider -x \
   -k "segmented_sort or score_ties or bbox_vote_vs or nms_indices" > gpurun_out/${T}_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/${T}_sanitizer_racecheck.log
tail -6 gpurun_out/${T}_sanitizer_racecheck.log
