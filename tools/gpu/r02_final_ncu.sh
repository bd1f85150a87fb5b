# Per-launch DRAM counters of the memory-bound kernels, final build (ReLU mask re-evaluated from y)
E="bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-kernel-table"
python $E > gpurun_out/r02_eager.json 2> gpurun_out/r02_eager.err && ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:act_fwd_kernel|bwd_apply|bwd_reduce|^fwd_kernel|^bwd_kernel|^wgrad_kernel|upsample|step_kernel" -s 52 -c 52 -f -o gpurun_out/r02_membound_final python $E > gpurun_out/r02_ncu3.log 2>&1
ls -la gpurun_out/r02_membound_final.ncu-rep
