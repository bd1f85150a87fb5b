# Final single-GPU records of round 2 (everything lands in gpurun_out/ -- keep it under 64 MiB -- the summaries are copied to profiles/ afterwards)
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -n 15 > gpurun_out/r02_gputests.log; tail -n 2 gpurun_out/r02_gputests.log; cp gpurun_out/parity_full_configs.json gpurun_out/r02_parity_full_configs.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_ref.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; cp gpurun_out/kernel_table.json gpurun_out/r02_kernel_table.json
python bench.py --model ico2ico_vae --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_vae_n1.json 2> gpurun_out/r02_vae.err; cp gpurun_out/kernel_table.json gpurun_out/r02_kernel_table_vae.json
python bench.py --level 6 --batch 16 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_i6_n1.json 2> gpurun_out/r02_i6.err; cp gpurun_out/kernel_table.json gpurun_out/r02_kernel_table_i6.json
python bench.py --steps 2000 --warmup 20 --no-cpu-baseline --no-kernel-table > gpurun_out/r02_bench_sustained_2000.json 2> gpurun_out/r02_sus.err
python tools/trace_step.py --out gpurun_out/r02_trace_n1.json > gpurun_out/r02_trace_n1.log 2>&1
python tools/trace_step.py --model ico2ico_vae --out gpurun_out/r02_trace_vae_n1.json > gpurun_out/r02_trace_vae_n1.log 2>&1
E="bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-table"
python $E > gpurun_out/r02_eager.json 2> gpurun_out/r02_eager.err && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 600 -c 520 --csv --log-file gpurun_out/r02_step_traffic.csv python $E > gpurun_out/r02_ncu1.log 2>&1
python $E > /dev/null 2>&1 && ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:act_fwd_kernel|bwd_apply|bwd_reduce|^fwd_kernel|^bwd_kernel|^wgrad_kernel|upsample|step_kernel" -s 52 -c 52 -f -o gpurun_out/r02_membound python $E > gpurun_out/r02_ncu2.log 2>&1
du -sh gpurun_out
for f in reference_arm n1 vae_n1 i6_n1 sustained_2000; do head -c 200 gpurun_out/r02_bench_$f.json; echo; done
