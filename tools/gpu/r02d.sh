python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_full.py 2>&1 | tail -25 > gpurun_out/r02d_gputests.log
python -m pytest tests/test_gpu_parity_full.py -q 2>&1 | tail -12 > gpurun_out/r02d_parity_full.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; cp gpurun_out/kernel_table.json gpurun_out/r02d_kernel_table.json
tail -n 6 gpurun_out/r02d_gputests.log; tail -n 6 gpurun_out/r02d_parity_full.log; head -c 400 gpurun_out/r02d_bench.json
