#!/bin/bash
# Final 1-GPU evidence of the round: tests (+ the unmodified reference driver when _refdata/ is pushed), the launch list
# and full captures of the shipped kernels, level tables, the bench line with BASELINE configs[4], the reference arm.
T=${TAG:-r02f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_gpu.txt 2>&1; nproc >> gpurun_out/${T}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -rA 2>&1 | grep -v "^PASSED" | tail -100 > gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
if [ -d _refdata/reference ]; then TAG=r02 bash tools/gpu_reference_driver_remote.sh; fi
timeout 600 python tools/level_conv_only.py 2048 5 > gpurun_out/${T}_level2048_conv_only.txt 2>&1
timeout 600 python tools/level_conv_only.py 1408 5 > gpurun_out/${T}_level1408_conv_only.txt 2>&1
tail -2 gpurun_out/${T}_level2048_conv_only.txt
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
   --clock-control none -k regex:conv_stream --csv --log-file gpurun_out/${T}_level2048_ncu.csv python tools/level_conv_only.py 2048 1 > /dev/null 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
   --log-file gpurun_out/launches_${T}.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_bench.log 2>&1
for spec in "conv4_2_fmt1 512 512 256 256 3 1 1" "conv2_2_fmt1 128 128 1024 1024 3 1 1" "conv1_2_fmt1 64 64 2048 2048 3 1 1" "conv4_2_fmt0 512 512 256 256 3 1 0"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_stream -s 2 -c 1 -f -o gpurun_out/${T}_$1 \
     python tools/ncu_one.py 8 $2 $3 $4 $5 $6 $7 $8 > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1_tc -s 2 -c 1 -f -o gpurun_out/${T}_conv1_tc \
   python tools/level_conv_only.py 2048 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:segmented_sort|mask_sweep|iou_mask|vote_reduce" -c 5 -f -o gpurun_out/${T}_post \
   python -c "import __graft_entry__ as g; g.smoke()" > /dev/null 2>&1
timeout 300 python tools/time_conv1.py 2048 1 1 > gpurun_out/${T}_conv1_tc_time.txt 2>&1; timeout 300 python tools/time_conv1.py 1408 16 1 >> gpurun_out/${T}_conv1_tc_time.txt 2>&1
timeout 300 python tools/plugin_breakdown.py 2>&1 | tail -7 > gpurun_out/${T}_plugin_breakdown.txt
timeout 300 python tools/resnet_level.py 1408 1 2>&1 | tail -16 > gpurun_out/${T}_resnet_level1408.txt
timeout 1200 python bench.py --steps 10 --warmup 3 --wider-shaped 3226 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err
tail -c 600 gpurun_out/${T}_bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
tail -c 300 gpurun_out/${T}_bench_reference.json
