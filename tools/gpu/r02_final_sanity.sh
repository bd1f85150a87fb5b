# Last call of the round: what the driver runs (smoke, default bench) plus the other two configurations and the timeline, final defaults
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -n 3 gpurun_out/r02_smoke.log
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; cp gpurun_out/kernel_table.json gpurun_out/r02_kernel_table.json
python bench.py --model ico2ico_vae --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table > gpurun_out/r02_bench_vae_n1.json 2> gpurun_out/r02_vae.err
python bench.py --level 6 --batch 16 --steps 20 --warmup 5 --no-cpu-baseline --no-kernel-table > gpurun_out/r02_bench_i6_n1.json 2> gpurun_out/r02_i6.err
python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --no-kernel-table > gpurun_out/r02_bench_sustained_1000.json 2> gpurun_out/r02_sus.err
python tools/trace_step.py --out gpurun_out/r02_trace_n1.json > gpurun_out/r02_trace_n1.log 2>&1
for f in n1 vae_n1 i6_n1 sustained_1000; do head -c 200 gpurun_out/r02_bench_$f.json; echo; done
