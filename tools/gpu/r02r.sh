# A/B: BatchNorm backward re-evaluating the ReLU mask from y (GIN_BN_MASK_FROM_Y=1) instead of reading it; GPU tests with it on
set -x
GIN_BN_MASK_FROM_Y=1 python -m pytest tests -m gpu -q 2>&1 | tail -n 30 | cut -c1-300 > gpurun_out/r02r_tests.log; tail -n 3 gpurun_out/r02r_tests.log
B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table"
GIN_BN_MASK_FROM_Y=0 python $B > gpurun_out/r02r_read.json 2> gpurun_out/r02r_read.err
GIN_BN_MASK_FROM_Y=1 python $B > gpurun_out/r02r_from_y.json 2> gpurun_out/r02r_from_y.err
GIN_BN_MASK_FROM_Y=1 python tools/trace_step.py --out gpurun_out/r02r_trace.json > gpurun_out/r02r_trace.log 2>&1
for f in read from_y; do head -c 220 gpurun_out/r02r_$f.json; echo; done
grep "bn::" gpurun_out/r02r_trace.log
