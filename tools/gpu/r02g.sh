python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_full.py 2>&1 | tail -25 > gpurun_out/r02g_gputests.log
python -m pytest tests/test_gpu_parity_full.py -q 2>&1 | tail -12 > gpurun_out/r02g_parity_full.log
cp gpurun_out/parity_full_configs.json gpurun_out/r02g_parity_full_configs.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; cp gpurun_out/kernel_table.json gpurun_out/r02g_kernel_table.json
GIN_Y_FP16=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table > gpurun_out/r02g_bench_y32.json 2> gpurun_out/r02g_bench2.err
tail -n 6 gpurun_out/r02g_gputests.log; tail -n 6 gpurun_out/r02g_parity_full.log; head -c 300 gpurun_out/r02g_bench.json; echo; head -c 300 gpurun_out/r02g_bench_y32.json
bash tools/gpu/r02f.sh
