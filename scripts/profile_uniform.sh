#!/bin/bash
# Profile pass for the uniform-run commit (reference benchmark shape, engine default = auto -> device): launch list of two
# cycles and a --set full capture of its own kernels.  Summaries: python scripts/summarize_uniform.py
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_uniform.csv \
    python scripts/prof_cycle.py reference auto > gpurun_out/r2_ncu_u0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:un_ -s 7 -c 7 -f -o gpurun_out/r2_uniform python scripts/prof_cycle.py reference auto > gpurun_out/r2_ncu_u1.log 2>&1
ls -la gpurun_out/r2_uniform.ncu-rep gpurun_out/r2_launches_uniform.csv
