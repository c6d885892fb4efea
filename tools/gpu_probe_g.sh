#!/bin/bash
# probes build: accumulator drain period (64-channel chunks per phase) on the fast-format layers
mkdir -p gpurun_out
for rep in 1 2; do
for g in 1 2 4; do
  echo "SHF_PROBE_G=$g rep $rep"
  SHF_PROBE_G=$g timeout 300 python tools/level_conv_only.py 2048 5 2>/dev/null | grep -A25 "operand format hf8" | grep -E "conv2_2|conv3_2|conv4_2|conv5_2|ALL"
done
done | tee gpurun_out/r02f_probe_g.txt
