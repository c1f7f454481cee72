set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 ) > gpurun_out/r2e_tests.log 2>&1
STEPS=10 timeout 400 tools/bench_workloads.sh million_triangles instanced_shard skinned_crowd stress_lit > gpurun_out/r2e_workloads.log 2>&1
cat gpurun_out/r2e_tests.log gpurun_out/r2e_workloads.log
