#!/bin/bash
# What the driver runs at round end, on the final code: the whole GPU suite, smoke(), the default bench line, the reference arm.
mkdir -p gpurun_out
T=${TAG:-r02z}
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -rA -s 2>&1 | grep -v "^PASSED" | tail -150 > gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_1gpu.json').read().strip().splitlines()[-1])
print('value %.2f ms %.2f frac %.4f e2e %.2f batched %.2f launches %s clk %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e_batched']['value'], d['gpu_launches'], d['clocks']))
PY
tail -3 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
tail -c 250 gpurun_out/${T}_bench_reference.json
