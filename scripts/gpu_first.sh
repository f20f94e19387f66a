#!/bin/bash
# first contact with the B200: smoke, gpu tests, quick timing
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
python __graft_entry__.py smoke 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python scripts/quick_time.py 2>&1 | tail -20
