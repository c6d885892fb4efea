#!/bin/bash
# small-level N tile heuristic: conv parity tests + whole-net parity, plugin breakdown A/B
mkdir -p gpurun_out
T=${TAG:-r02i}
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_net.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -5 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
for v in 1 0 1 0; do echo "SHF_CONV_SMALL_BN64=$v"; SHF_CONV_SMALL_BN64=$v timeout 300 python tools/plugin_breakdown.py 2>&1 | tail -7; done | tee gpurun_out/${T}_plugin_breakdown.txt
