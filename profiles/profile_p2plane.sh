#!/bin/bash
# one gpurun call: the point-to-plane SE(3) solve (north-star extension) at the headline size -- device times per execution
# path, the whole frame with that term, and one full ncu capture of the cooperative kernel
set -x
mkdir -p gpurun_out
timeout 200 python tools/bench_p2plane.py > gpurun_out/r01j_p2plane.json 2> gpurun_out/r01j_p2plane.err
cat gpurun_out/r01j_p2plane.json; grep cycles gpurun_out/r01j_p2plane.err | head -2
timeout 300 python tools/bench_configs.py C3p > gpurun_out/r01j_c3p.json 2> gpurun_out/r01j_c3p.err
cat gpurun_out/r01j_c3p.json
DFU_P2P_PROFILE_ONCE=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:kp_persistent -s 12 -c 1 -f \
    -o gpurun_out/r01j_p2plane python tools/bench_p2plane.py > gpurun_out/r01j_ncu_p2plane.log 2>&1
tail -3 gpurun_out/r01j_ncu_p2plane.log
ls -la gpurun_out | tail -5
