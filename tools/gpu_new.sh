#!/bin/bash
# the whole GPU suite (incl. the tests added in this session: ResNet ops + net, native driver, WIDER evaluator)
mkdir -p gpurun_out
T=${TAG:-r02k}
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -rA -s 2>&1 | grep -v "^PASSED" | tail -150 > gpurun_out/${T}_pytest.log
grep -E "resnet|passed|failed|FAILED|Error" gpurun_out/${T}_pytest.log | tail -40
