GIN_PAIR=2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_pair2.json 2> gpurun_out/r02e_1.err; cp gpurun_out/kernel_table.json gpurun_out/r02e_kernel_table_pair2.json
GIN_PAIR=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_pair0.json 2> gpurun_out/r02e_2.err; cp gpurun_out/kernel_table.json gpurun_out/r02e_kernel_table_pair0.json
head -c 300 gpurun_out/r02e_bench_pair2.json; echo; head -c 300 gpurun_out/r02e_bench_pair0.json
