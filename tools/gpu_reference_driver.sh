#!/bin/bash
# Run HERE (the container with /root/reference): pushes a copy of the reference tree as git-ignored DATA, runs the
# unmodified-driver test and a timed 1024x1024 run of the reference's own train_test.py on the B200, removes the copy.
set -e
cd "$(dirname "$0")/.."
mkdir -p _refdata
rm -rf _refdata/reference
cp -r /root/reference _refdata/reference
chmod -R u+w _refdata/reference
trap 'rm -rf _refdata' EXIT
/usr/local/graft/bin/gpurun --timeout ${SHF_GPURUN_TIMEOUT:-1500} -- "${SHF_GPURUN_CMD:-bash tools/gpu_reference_driver_remote.sh}" 2>&1 | tail -40
