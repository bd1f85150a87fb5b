# Final multi-GPU records (one 8-GPU box): DP parity through the captured step at 2/4/8 ranks, AE / VAE / I6 scaling lines
B="--steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table"
python -m pytest tests/test_gpu_dp.py -q 2>&1 | tail -5 > gpurun_out/r02_dp_tests.log; tail -n 3 gpurun_out/r02_dp_tests.log
cat gpurun_out/dp_parity_n2.json gpurun_out/dp_parity_n4.json gpurun_out/dp_parity_n8.json
port=29600
run() { out=$1; n=$2; shift 2; port=$((port+1)); if [ $n = 1 ]; then python bench.py $B "$@" > gpurun_out/$out 2> gpurun_out/$out.err; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n $B "$@" > gpurun_out/$out 2> gpurun_out/$out.err; fi; head -c 230 gpurun_out/$out; echo; }
for n in 1 2 4 8; do run r02_bench_n$n.scale.json $n; done
run r02_bench_vae_n8.json 8 --model ico2ico_vae
run r02_bench_i6_n8.json 8 --level 6 --batch 16
