python -m pytest tests/test_gpu_dp.py -q -x 2>&1 | tail -8 > gpurun_out/r02i_dp.log
tail -n 4 gpurun_out/r02i_dp.log; cat gpurun_out/dp_parity_n2.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table > gpurun_out/r02i_n1.json 2> gpurun_out/r02i_n1.err
for v in "8" "64" "2"; do
GIN_DP_BUCKET_MB=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table > gpurun_out/r02i_n2_b$v.json 2> gpurun_out/r02i_n2_b$v.err
done
NCCL_MAX_CTAS=4 GIN_DP_BUCKET_MB=8 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table > gpurun_out/r02i_n2_b8_cta4.json 2> gpurun_out/r02i_n2_b8_cta4.err
for f in gpurun_out/r02i_n1.json gpurun_out/r02i_n2_b8.json gpurun_out/r02i_n2_b64.json gpurun_out/r02i_n2_b2.json gpurun_out/r02i_n2_b8_cta4.json; do echo $f; head -c 230 $f; echo; done
