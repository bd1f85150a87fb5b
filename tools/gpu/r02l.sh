B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 $B --gpus 2 > gpurun_out/r02l_n2.json 2>/dev/null
GIN_DP_NOCOMM=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 $B --gpus 2 > gpurun_out/r02l_nocomm.json 2>/dev/null
GIN_DP_BUCKET_MB=64 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 $B --gpus 2 > gpurun_out/r02l_n2_one.json 2>/dev/null
python $B > gpurun_out/r02l_n1.json 2>/dev/null
GIN_NTILE_MAX=64 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_nt64.json 2>/dev/null; cp gpurun_out/kernel_table.json gpurun_out/r02l_kernel_table_nt64.json
GIN_NTILE_MAX=128 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_nt128.json 2>/dev/null; cp gpurun_out/kernel_table.json gpurun_out/r02l_kernel_table_nt128.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_nt256.json 2>/dev/null; cp gpurun_out/kernel_table.json gpurun_out/r02l_kernel_table_nt256.json
for f in n1 n2 nocomm n2_one nt64 nt128 nt256; do python - <<PY
import json
d=json.loads(open('gpurun_out/r02l_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['ms_per_step'],4), d.get('ms_per_step_per_rank'))
PY
done
