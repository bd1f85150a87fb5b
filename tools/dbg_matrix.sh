#!/bin/bash
for spec in "128 64 1 5 36 dgrad" "256 128 1 4 36 dgrad" "256 256 1 3 36 dgrad" "64 64 1 5 36 dgrad" "128 64 1 5 36 wgrad" "256 256 1 3 36 wgrad"; do
  python tools/run_layer.py $spec 20 | tail -1
done
