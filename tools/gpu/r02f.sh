bash tools/prof_layers.sh r02f "128 128 1 4 36 fwd" "128 128 1 4 36 dgrad" "128 128 1 4 36 wgrad" "256 256 1 4 36 fwd" > gpurun_out/r02f_prof.log 2>&1
python tools/layer_sweep.py gpurun_out/r02_layer_sweep.md > gpurun_out/r02f_sweep.log 2>&1
tail -5 gpurun_out/r02f_prof.log; tail -30 gpurun_out/r02f_sweep.log
