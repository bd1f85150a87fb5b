# A/B of the shared-memory-constant BatchNorm kernels and the one-launch Adam (both opt-in here), after the GPU tests with both on
set -x
GIN_BN_SMEM=1 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 25 > gpurun_out/r02q_tests.log; tail -n 3 gpurun_out/r02q_tests.log
B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table"
GIN_BN_SMEM=0 python $B --optimizer torch > gpurun_out/r02q_base.json 2> gpurun_out/r02q_base.err
GIN_BN_SMEM=1 python $B --optimizer torch > gpurun_out/r02q_bn.json 2> gpurun_out/r02q_bn.err
GIN_BN_SMEM=1 python $B --optimizer gin > gpurun_out/r02q_bn_adam.json 2> gpurun_out/r02q_bn_adam.err
GIN_BN_SMEM=1 GIN_BENCH_OPTIMIZER=gin python tools/trace_step.py --out gpurun_out/r02q_trace.json > gpurun_out/r02q_trace.log 2>&1
for f in base bn bn_adam; do head -c 220 gpurun_out/r02q_$f.json; echo; tail -n 2 gpurun_out/r02q_$f.err; done
head -n 45 gpurun_out/r02q_trace.log
