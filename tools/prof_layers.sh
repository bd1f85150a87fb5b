#!/bin/bash
# ncu --set full captures of the tcgen05 conv kernels on real layer shapes (one GPU). Usage: tools/prof_layers.sh <tag> "<cin cout stride level B pass>" ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for spec in "$@"; do
  set -- $spec
  name="${tag}_$6_$1x$2_s$3_L$4"
  python tools/run_layer.py $spec 3 > gpurun_out/${name}.plain.log 2>&1 || { echo "plain run failed: $spec"; tail -5 gpurun_out/${name}.plain.log; continue; }
  ncu --set full --clock-control none --import-source on -k "regex:patch_conv|wgrad_patch_kernel|gather_gemm_tc_kernel|wgrad_tc_kernel" -s 2 -c 1 -f -o gpurun_out/${name} python tools/run_layer.py $spec 1 > gpurun_out/${name}.ncu.log 2>&1
  tail -1 gpurun_out/${name}.plain.log
done
