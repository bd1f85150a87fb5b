#!/bin/bash
for spec in "256 256 1 4 36 fwd" "128 128 1 5 36 fwd" "256 128 1 4 36 fwd" "256 512 1 3 36 fwd" "128 512 2 4 36 fwd" "256 256 1 4 36 dgrad" "128 128 1 5 36 dgrad"; do
  python tools/run_layer.py $spec 20 | tail -1
  GIN_PAIR=0 python tools/run_layer.py $spec 20 | tail -1
done
