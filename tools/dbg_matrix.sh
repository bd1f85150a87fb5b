#!/bin/bash
for spec in "128 64 1 5 36 fwd" "64 64 1 5 36 fwd" "128 128 1 4 36 fwd" "256 128 1 4 36 fwd" "256 256 1 3 36 fwd" "256 256 1 2 36 fwd"; do
  python tools/run_layer.py $spec 20 | tail -1
done
for d in 1 2 4 3 5 6 7; do echo -n "dbg=$d  "; GIN_DBG=$d python tools/run_layer.py 128 64 1 5 36 fwd 20 | tail -1; done
