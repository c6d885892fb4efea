#!/bin/bash
# the Python 3 CLI under torchrun on 2 GPUs (NCCL): same detections as the single-process run
mkdir -p gpurun_out
W=/tmp/shf_cli2
rm -rf $W; mkdir -p $W
python - <<PY
import os, sys, cv2, numpy as np
sys.path.insert(0, '.')
from smallhardface_b200 import config as C, deploy
C.write_builtin_tree('$W/tree')
print(deploy.write_synthetic_deployment('$W/deploy', dilation=True)[1])
os.makedirs('$W/imgs/a', exist_ok=True)
rng = np.random.RandomState(1)
for i, (h, w) in enumerate([(96, 128), (128, 96), (96, 128), (80, 80), (128, 96), (96, 128), (80, 80)]):
    cv2.imwrite('$W/imgs/a/im%d.png' % i, rng.randint(0, 256, (h, w, 3)).astype(np.uint8))
PY
PRE="--root $W/tree --conf configs/smallhardface.toml --batch 2"
ARGS="--amend DATA_DIR $W/imgs TEST.DB general_png TEST.MODEL $W/deploy/synthetic_dil_seed3.caffemodel TEST.SCALES [100,300] TEST.NO_CACHE False"
timeout 600 python -m smallhardface_b200.run_test $PRE --output $W/out1 $ARGS > $W/log1 2>&1 || tail -5 $W/log1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 -m smallhardface_b200.run_test $PRE --output $W/out2 $ARGS > $W/log2 2>&1 || tail -8 $W/log2
python - <<PY | tee gpurun_out/r02f_cli_2gpu.txt
import glob, pickle, numpy as np
a = pickle.load(open(glob.glob('$W/out1/face/general_png/*/detections.pkl')[0], 'rb'))
dirs = glob.glob('$W/out2/face/general_png/*')
b = pickle.load(open(glob.glob('$W/out2/face/general_png/*/detections.pkl')[0], 'rb'))
same = all(x.shape == y.shape and np.array_equal(x, y) for x, y in zip(a[1], b[1]))
print("native CLI: 1 process vs torchrun x2 (NCCL): %d images, %d detections, output dirs of the 2-rank job: %d, identical results: %s"
      % (len(a[1]), sum(len(x) for x in a[1]), len(dirs), same))
PY
