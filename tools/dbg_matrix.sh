#!/bin/bash
for spec in "256 128 1 4 36 fwd" "128 64 1 5 36 fwd" "128 128 1 4 36 fwd" "64 64 1 5 36 fwd" "256 256 1 3 36 fwd" "64 128 2 5 36 fwd" "64 128 2 5 36 dgrad" "128 128 1 5 36 fwd" "64 256 2 5 36 fwd"; do
  python tools/run_layer.py $spec 20 | tail -1
done
