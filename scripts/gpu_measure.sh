#!/bin/bash
# round-1 measurement pass: gpu tests, bench (both arms), ncu launch list of the bench command, full capture of the sweep
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours.json
python bench.py --steps 5 --warmup 3 --masks 2>&1 | tail -1 | tee gpurun_out/bench_ours_masks.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:yk_sweep -s 30 -c 3 -o gpurun_out/sweep_r1 \
    python scripts/prof_cycle.py > gpurun_out/ncu_sweep.log 2>&1
tail -3 gpurun_out/ncu_sweep.log
ls -la gpurun_out
