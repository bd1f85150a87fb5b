# Final build on a 2-GPU box: N=1 and N=2 bench lines
set -x
B="--steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table"
python bench.py $B > gpurun_out/r02_final2_n1.json 2> gpurun_out/r02_final2_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 $B > gpurun_out/r02_final2_n2.json 2> gpurun_out/r02_final2_n2.err
for f in n1 n2; do head -c 230 gpurun_out/r02_final2_$f.json; echo; done
