#!/bin/bash
# one gpurun call: bench, launch list, full captures of the two dominant kernels
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r01h_bench.json 2> gpurun_out/r01h_bench.err
tail -c 600 gpurun_out/r01h_bench.err
# steady state: 30 warm-up frame pairs first (the field needs a few bend cycles to become periodic), then the last ~8 frames
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 820 -c 120 --csv --log-file gpurun_out/r01h_launches.csv \
    python bench.py --steps 3 --warmup 30 --no-cpu-baseline --no-variant > gpurun_out/r01h_under_ncu.log 2>&1
# steady-state frames: skip the first launches (cache fill) -- the 6th..8th integrate launches, 4th..5th solver launches
timeout 600 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -s 120 -c 4 -f -o gpurun_out/r01h_integrate \
    python bench.py --steps 3 --warmup 30 --no-cpu-baseline --no-variant > gpurun_out/r01h_ncu_integrate.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_persistent -s 60 -c 2 -f -o gpurun_out/r01h_solver \
    python bench.py --steps 3 --warmup 30 --no-cpu-baseline --no-variant > gpurun_out/r01h_ncu_solver.log 2>&1
ls -la gpurun_out
