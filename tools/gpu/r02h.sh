timeout 600 python -m pytest tests/test_gpu_layers.py -m gpu -q -x -k "tc_matches or alternative" 2>&1 | tail -15 > gpurun_out/r02h_tc.log
tail -n 5 gpurun_out/r02h_tc.log
for cl in 1 2 4; do
GIN_CLUSTER=$cl timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_cl$cl.json 2> gpurun_out/r02h_cl$cl.err; cp gpurun_out/kernel_table.json gpurun_out/r02h_kernel_table_cl$cl.json
head -c 250 gpurun_out/r02h_bench_cl$cl.json; echo
done
