python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/diag/dp_parity_worker.py ico2ico 6 > gpurun_out/r02k_worker.log 2>&1
grep -v "Warning\|warn\|_engine_run\|execution_engine" gpurun_out/r02k_worker.log | tail -5
B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table"
CUDA_VISIBLE_DEVICES=0 python $B > gpurun_out/r02k_gpu0.json 2>/dev/null
CUDA_VISIBLE_DEVICES=1 python $B > gpurun_out/r02k_gpu1.json 2>/dev/null
CUDA_VISIBLE_DEVICES=0 python $B > gpurun_out/r02k_both0.json 2>/dev/null &
CUDA_VISIBLE_DEVICES=1 python $B > gpurun_out/r02k_both1.json 2>/dev/null
wait
GIN_DP_NOCOMM=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 $B --gpus 2 > gpurun_out/r02k_nocomm.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 $B --gpus 2 > gpurun_out/r02k_n2.json 2>/dev/null
for f in gpu0 gpu1 both0 both1 nocomm n2; do python - <<PY
import json
d=json.loads(open('gpurun_out/r02k_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['ms_per_step'],4), d.get('ms_per_step_per_rank'), d['clocks'])
PY
done
