#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/quick_time.py 2>&1 | tail -6
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference_r1e.json
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_r1e.json
python bench.py --steps 10 --warmup 3 --masks 2>&1 | tail -1 > gpurun_out/bench_ours_masks_r1e.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_r1e.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
YK_BATCH=4096 ncu --set full --clock-control none --import-source on -k regex:yk_sweep -s 16 -c 3 -o gpurun_out/sweep_r1e python scripts/prof_cycle.py > gpurun_out/ncu_sweep.log 2>&1
YK_BATCH=4096 ncu --set full --clock-control none --import-source on -k regex:yk_sweep -s 16 -c 3 -o gpurun_out/sweep_r1e_masks python scripts/prof_cycle.py --masks > gpurun_out/ncu_sweep_masks.log 2>&1
