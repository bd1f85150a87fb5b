B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table"
python $B > gpurun_out/r02p_n1.json 2>gpurun_out/r02p_n1.err
GIN_DP_SPARE_SMS=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 $B --gpus 2 > gpurun_out/r02p_n2.json 2> gpurun_out/r02p_n2.err
GIN_DP_SPARE_SMS=0 GIN_DP_BUCKET_MB=4 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 $B --gpus 2 > gpurun_out/r02p_n2_b4.json 2> gpurun_out/r02p_n2_b4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/trace_step.py --out gpurun_out/r02_trace_n2.json > gpurun_out/r02p_trace_n2.log 2>&1; grep "replay\|NCCL at" gpurun_out/r02p_trace_n2.log | tail -6
python -m pytest tests/test_gpu_dp.py -q 2>&1 | tail -3
for f in n1 n2 n2_b4; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02p_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['ms_per_step'],4), d.get('ms_per_step_per_rank'))
except Exception as e: print('$f ERR', e)
PY
done
tail -3 gpurun_out/r02p_n2.err
