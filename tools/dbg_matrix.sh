#!/bin/bash
for spec in "64 128 2 5 36" "128 256 2 4 36" "256 256 2 3 36"; do
  for pass in fwd dgrad wgrad; do python tools/run_layer.py $spec $pass 20 | tail -1; done
done
for spec in "128 64 1 5 36" "256 256 1 3 36"; do
  for pass in fwd dgrad wgrad; do python tools/run_layer.py $spec $pass 20 | tail -1; done
done
