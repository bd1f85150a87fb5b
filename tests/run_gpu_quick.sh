#!/bin/bash
# one GPU-box call: parity tests, then whatever else is passed
set -o pipefail
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
