python -m pytest tests/test_gpu_layers.py tests/test_gpu_models.py tests/test_gpu_parity_full.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02m_tests.log; tail -n 3 gpurun_out/r02m_tests.log
B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table"
python $B > gpurun_out/r02m_n1.json 2>/dev/null
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 $B --gpus 2 > gpurun_out/r02m_$tag.json 2> gpurun_out/r02m_$tag.err; }
run spare0 GIN_DP_SPARE_SMS=0
run spare4 GIN_DP_SPARE_SMS=4
run spare8 GIN_DP_SPARE_SMS=8
run spare16 GIN_DP_SPARE_SMS=16
run spare8_b2 GIN_DP_SPARE_SMS=8 GIN_DP_BUCKET_MB=2
for f in n1 spare0 spare4 spare8 spare16 spare8_b2; do python - <<PY
import json
d=json.loads(open('gpurun_out/r02m_$f.json').read().strip().splitlines()[-1])
print('$f', round(d['ms_per_step'],4), d.get('ms_per_step_per_rank'))
PY
done
