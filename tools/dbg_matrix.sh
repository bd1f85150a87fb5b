#!/bin/bash
for spec in "128 128 1 5 36 dgrad" "256 256 1 4 36 dgrad" "256 256 1 3 36 dgrad" "64 256 2 5 36 dgrad" "64 64 1 5 36 dgrad"; do
  python tools/run_layer.py $spec 20 | tail -1
  GIN_SEAM=gather python tools/run_layer.py $spec 20 | tail -1
done
