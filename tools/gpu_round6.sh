#!/bin/bash
# r02f: MMA issue-loop restructure (format template, resident-weights fast path) -- tests, level table, bench
T=${TAG:-r02f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -rA 2>&1 | grep -v "^PASSED" | tail -60 > gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
timeout 600 python tools/level_conv_only.py 2048 5 > gpurun_out/${T}_level2048_conv_only.txt 2>&1
grep -E "conv1_2|conv2_1|conv2_2|conv3_3|conv4_2|conv5_2|head_1|ALL" gpurun_out/${T}_level2048_conv_only.txt
timeout 1200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_1gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'batched', d['e2e_batched']['value'], d['clocks'])
print(d['e2e']['breakdown_ms_per_image'])
PY
