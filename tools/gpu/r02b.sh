python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_full.py 2>&1 | tail -25 > gpurun_out/r02b_gputests.log
python -m pytest tests/test_gpu_parity_full.py -q 2>&1 | tail -40 > gpurun_out/r02b_parity_full.log
cp gpurun_out/parity_full_configs.json gpurun_out/r02b_parity_full_configs.json
python tests/diag/precision_study.py --out gpurun_out/r02b_precision_study.json > gpurun_out/r02b_precision.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
tail -n 6 gpurun_out/r02b_gputests.log; tail -n 8 gpurun_out/r02b_parity_full.log; grep -v Warning gpurun_out/r02b_precision.log | tail -n 22; head -c 600 gpurun_out/r02b_bench.json
