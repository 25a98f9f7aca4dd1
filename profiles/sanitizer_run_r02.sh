#!/bin/bash
# compute-sanitizer over the round's kernels (small configurations; cooperative launches and clusters are supported by the tool)
mkdir -p gpurun_out
SEL1='solver_every_execution_path or solver_fixed_iteration or self_contained or lanes_per_query or tsdf_dense_depth or tsdf_warped_bit_exact or p2plane_solver_matches'
SEL2='marching_cubes_bit_exact or dfu_frame or overlapped_frame'
{
echo "== memcheck: tests/test_gpu_parity.py -k \"$SEL1\""
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 97 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "$SEL1" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds|leaked" | tail -8
echo "== memcheck: marching cubes / dfu_frame / overlapped schedule"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 97 python -m pytest tests/test_marching_cubes.py tests/test_gpu_frontend.py -q -m gpu -x -k "$SEL2" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" | tail -8
echo "== racecheck: solver 3r / version 4 / generic, dense integrate"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 97 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "solver_every_execution_path and (p3 or p4) or tsdf_dense_depth" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -8
echo "== racecheck: marching cubes"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 97 python -m pytest tests/test_marching_cubes.py -q -m gpu -x -k "marching_cubes_bit_exact" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -8
echo "== synccheck: solver 3r / version 4"
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 97 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "solver_every_execution_path and (p3 or p4)" 2>&1 | grep -E "passed|failed|SYNCCHECK SUMMARY|ERROR SUMMARY|divergent" | tail -8
} > gpurun_out/r02_sanitizer.txt 2>&1
cat gpurun_out/r02_sanitizer.txt
