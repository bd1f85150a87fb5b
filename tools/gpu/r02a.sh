python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_full.py 2>&1 | tail -15 > gpurun_out/r02a_gputests.log
python -m pytest tests/test_gpu_parity_full.py -q 2>&1 | tail -40 > gpurun_out/r02a_parity_full.log
python tests/diag/precision_study.py > gpurun_out/r02a_precision.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; cp gpurun_out/kernel_table.json gpurun_out/r02a_kernel_table.json
timeout 60 tools/exp/umma_peak_test 4000 1 > gpurun_out/r02a_umma_peak.log 2>&1
python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-table > gpurun_out/r02a_eager.json 2>gpurun_out/r02a_eager.err && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 700 -c 450 --csv --log-file gpurun_out/r02a_step_traffic.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-table > gpurun_out/r02a_ncu.log 2>&1
timeout 30 tools/exp/umma_peak_test 2000 2 > gpurun_out/r02a_umma_pair.log 2>&1
tail -3 gpurun_out/r02a_gputests.log gpurun_out/r02a_parity_full.log; tail -12 gpurun_out/r02a_precision.log; cat gpurun_out/r02a_umma_peak.log gpurun_out/r02a_umma_pair.log; head -c 1500 gpurun_out/r02a_bench.json
