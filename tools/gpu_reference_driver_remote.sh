#!/bin/bash
# GPU-box half of tools/gpu_reference_driver.sh
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_reference_driver.py -m gpu -q -p no:cacheprovider -rA -s 2>&1 | tail -30 > gpurun_out/${TAG:-r02}_reference_driver_test.log
tail -5 gpurun_out/${TAG:-r02}_reference_driver_test.log
# a timed run of the reference's driver on bench-sized images (its own detect / misc timers, lib/test.py:244-256)
python - <<'PY'
import os, sys, cv2
sys.path.insert(0, '.')
from smallhardface_b200 import deploy
os.makedirs('/tmp/shf_imgs1024', exist_ok=True)
for i in range(8):
    cv2.imwrite('/tmp/shf_imgs1024/im%d.png' % i, deploy.synthetic_image(3 + i, (1024, 1024)))
print(deploy.write_synthetic_deployment('/tmp/shf_b200_deploy', dilation=True)[1])
PY
rm -rf /tmp/shf_ref_run && cp -r _refdata/reference /tmp/shf_ref_run
( time timeout 1200 python tools/run_reference_driver.py --reference-root /tmp/shf_ref_run -- --train false --conf configs/smallhardface.toml \
    --amend DATA_DIR /tmp/shf_imgs1024 TEST.DB general_png TEST.MODEL /tmp/shf_b200_deploy/synthetic_dil_seed3.caffemodel TEST.GPU_ID "[0]" ) \
    > gpurun_out/${TAG:-r02}_reference_driver_1024.log 2>&1
cat /tmp/shf_ref_run/output/face/general_png/*/stderr.log | tail -15 >> gpurun_out/${TAG:-r02}_reference_driver_1024.log
tr '\r' '\n' < gpurun_out/${TAG:-r02}_reference_driver_1024.log | tail -12
