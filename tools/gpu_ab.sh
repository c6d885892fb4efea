#!/bin/bash
# same-box A/B of the MMA-issue grouping (SHF_CONV_GROUP_ROWS), three alternating repetitions each
mkdir -p gpurun_out
for i in 1 2 3; do
  for g in 1 0; do
    SHF_CONV_GROUP_ROWS=$g timeout 600 python tools/level_conv_only.py 2048 8 2>/dev/null | grep -E "ALL|conv2_2|conv4_2|conv2_1" | sed "s/^/group_rows=$g rep$i /" >> gpurun_out/r02_ab_group_rows.txt
  done
done
for g in 1 0 1 0; do
  SHF_CONV_GROUP_ROWS=$g timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('group_rows=$g bench value %.2f ms %.2f frac %.4f clk %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz']))" >> gpurun_out/r02_ab_group_rows.txt
done
cat gpurun_out/r02_ab_group_rows.txt
