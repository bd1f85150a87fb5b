# Data-parallel check with the final defaults on a 2-GPU box: DP parity test through the captured step, N=1 and N=2 bench lines, in-graph timeline at N=2
set -x
B="--steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table"
python -m pytest tests -m gpu -q 2>&1 | tail -n 15 > gpurun_out/r02_gputests.log; tail -n 2 gpurun_out/r02_gputests.log; cp gpurun_out/dp_parity_n2.json gpurun_out/r02_dp_parity_n2.json
python bench.py $B > gpurun_out/r02_n2box_n1.json 2> gpurun_out/r02_n2box_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 $B > gpurun_out/r02_n2box_n2.json 2> gpurun_out/r02_n2box_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/trace_step.py --out gpurun_out/r02_trace_n2.json > gpurun_out/r02_trace_n2.log 2>&1
for f in n1 n2; do head -c 230 gpurun_out/r02_n2box_$f.json; echo; tail -n 2 gpurun_out/r02_n2box_$f.err | cut -c1-300; done
head -n 30 gpurun_out/r02_trace_n2.log
