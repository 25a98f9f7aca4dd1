#!/bin/bash
# Round 2, one gpurun call: bench line, steady-state launch list, full captures of the three dominant kernels + the dense integrate
# (raw .ncu-rep files stay in gpurun_out/; profiles/make_summaries_r02.sh turns them into the tracked summaries)
set -x
mkdir -p gpurun_out
T=r02j
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 400 gpurun_out/${T}_bench.err
# steady state: 30 warm-up frame pairs first (the field needs a few bend cycles to become periodic), then the last frames
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 690 -c 120 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 3 --warmup 30 --loops 1 --no-cpu-baseline --no-variant --no-overlap > gpurun_out/${T}_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -s 130 -c 4 -f -o gpurun_out/${T}_integrate \
    python bench.py --steps 3 --warmup 30 --loops 1 --no-cpu-baseline --no-variant --no-overlap > gpurun_out/${T}_ncu_integrate.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_persistent -s 60 -c 2 -f -o gpurun_out/${T}_solver \
    python bench.py --steps 3 --warmup 30 --loops 1 --no-cpu-baseline --no-variant --no-overlap > gpurun_out/${T}_ncu_solver.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:points_grid -s 60 -c 2 -f -o gpurun_out/${T}_knn \
    python bench.py --steps 3 --warmup 30 --loops 1 --no-cpu-baseline --no-variant --no-overlap > gpurun_out/${T}_ncu_knn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -s 40 -c 2 -f -o gpurun_out/${T}_dense \
    python tools/dense_integrate.py > gpurun_out/${T}_ncu_dense.log 2>&1
ls -la gpurun_out | tail -12
