#!/bin/bash
# One measurement pass on the B200 box (run through gpurun): gpu tests, quick timings, both bench arms, the ncu launch
# list of the bench command and the --set full capture of the sweep kernel.  Usage: bash scripts/measure_round.sh r2a
# Afterwards, here:  python scripts/summarize_profiles.py r2a launches_bench_r2a.csv sweep_r2a.ncu-rep
TAG=${1:-rX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/time_policies.py 2>&1 | tail -4
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference_${TAG}.json
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours_${TAG}.json
python bench.py --steps 10 --warmup 3 --masks 2>&1 | tail -1 > gpurun_out/bench_ours_masks_${TAG}.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# full captures of the sweep kernel at full load (one row per ask); FULL=0 skips them when the kernel has not changed
if [ "${FULL:-1}" = "1" ]; then
YK_NO_ROW_SHARING=1 YK_BATCH=4096 ncu --set full --clock-control none --import-source on -k regex:yk_sweep -s 16 -c 3 -o gpurun_out/sweep_${TAG} python scripts/prof_cycle.py > gpurun_out/ncu_sweep.log 2>&1
YK_BATCH=4096 ncu --set full --clock-control none --import-source on -k regex:yk_sweep -s 16 -c 3 -o gpurun_out/sweep_${TAG}_masks python scripts/prof_cycle.py --masks > gpurun_out/ncu_sweep_masks.log 2>&1
fi
