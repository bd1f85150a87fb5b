python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/diag/dp_parity_worker.py ico2ico 6 > gpurun_out/r02j_worker.log 2>&1
grep -v "Warning\|warn" gpurun_out/r02j_worker.log | tail -30
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table > gpurun_out/r02j_$tag.json 2> gpurun_out/r02j_$tag.err; echo $tag; head -c 200 gpurun_out/r02j_$tag.json; echo; }
run early1 GIN_DP_EARLY=1
run early0 GIN_DP_EARLY=0
run nocomm GIN_DP_NOCOMM=1
run early0_one GIN_DP_EARLY=0 GIN_DP_BUCKET_MB=64
