set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2d_tests.log 2>&1
STEPS=10 timeout 400 tools/bench_workloads.sh million_triangles instanced_shard skinned_crowd skinned_crowd_animated stress_lit million_one_part > gpurun_out/r2d_workloads.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_vertex -c 2 -o gpurun_out/r2d_vertex python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2d_ncu.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_vertex -c 2 -o gpurun_out/r2d_vertex_inst python bench.py --steps 2 --warmup 1 --no-cpu-baseline --workload instanced_shard > gpurun_out/r2d_ncu2.log 2>&1
cat gpurun_out/r2d_tests.log gpurun_out/r2d_workloads.log
