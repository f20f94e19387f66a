#!/bin/bash
# Round-2 profile pass on the B200 box (through gpurun): the ncu launch list of the bench's headline arm and --set full
# captures of the sweep kernel in its three launch shapes + the lattice kernel.  Summarise here afterwards with
#   python scripts/summarize_round2.py
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --quick --steps 2 --warmup 3 > gpurun_out/r2_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:yk_sweep -s 1 -c 2 -f -o gpurun_out/r2_sweep_epochrows python scripts/prof_cycle.py config2 > gpurun_out/r2_ncu_a.log 2>&1
YK_NO_ROW_SHARING=1 ncu --set full --clock-control none --import-source on -k regex:yk_sweep -s 4 -c 2 -f -o gpurun_out/r2_sweep_fullload python scripts/prof_cycle.py config2 > gpurun_out/r2_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:yk_sweep -s 4 -c 2 -f -o gpurun_out/r2_sweep_masks python scripts/prof_cycle.py config3 > gpurun_out/r2_ncu_c.log 2>&1
if [ "${LATTICE:-1}" = "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:yk_lattice_kernel -c 1 -f -o gpurun_out/r2_lattice_cfg2 python scripts/prof_cycle.py config2 device > gpurun_out/r2_ncu_d.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep
