#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/quick_time.py 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_r1b.json
ncu --set full --clock-control none --import-source on -k regex:yk_sweep -s 30 -c 3 -o gpurun_out/sweep_r1b python scripts/prof_cycle.py > gpurun_out/ncu_sweep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:yk_sweep -s 30 -c 3 -o gpurun_out/sweep_r1b_masks python scripts/prof_cycle.py --masks > gpurun_out/ncu_sweep_masks.log 2>&1
