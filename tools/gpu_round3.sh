#!/bin/bash
# r02c: epilogue rewrite check -- tests, the unmodified reference driver, level table, bench line
T=r02c
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -rA 2>&1 | grep -v "^PASSED" | tail -120 > gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
if [ -d _refdata/reference ]; then bash tools/gpu_reference_driver_remote.sh; fi
timeout 600 python tools/level_conv_only.py 2048 5 > gpurun_out/${T}_level2048_conv_only.txt 2>&1
grep -E "conv1_2|conv2_1|conv2_2|conv4_2|ALL" gpurun_out/${T}_level2048_conv_only.txt
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02c_bench_1gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'batched', d['e2e_batched']['value'], d['clocks'])
PY
