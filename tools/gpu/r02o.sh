python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02o_tests.log; tail -n 3 gpurun_out/r02o_tests.log
python tools/trace_step.py --out gpurun_out/r02_trace_n1.json > gpurun_out/r02o_trace_n1.log 2>&1; grep "replay\|pack_weights\|stats_final" gpurun_out/r02o_trace_n1.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/trace_step.py --out gpurun_out/r02_trace_n2.json > gpurun_out/r02o_trace_n2.log 2>&1; grep "replay\|NCCL at\|at::native" gpurun_out/r02o_trace_n2.log | tail -8
B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table"
python $B > gpurun_out/r02o_n1.json 2>/dev/null
GIN_DP_SPARE_SMS=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 $B --gpus 2 > gpurun_out/r02o_n2.json 2> gpurun_out/r02o_n2.err
for f in n1 n2; do python - <<PY
import json
d=json.loads(open('gpurun_out/r02o_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['ms_per_step'],4), d.get('ms_per_step_per_rank'))
PY
done
