set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2k_tests.log 2>&1
STEPS=10 timeout 600 tools/bench_workloads.sh million_triangles stress_lit instanced_shard skinned_crowd million_one_part > gpurun_out/r2k_workloads.log 2>&1
echo "---- T3D_FUSE_SORT=0" >> gpurun_out/r2k_workloads.log
T3D_FUSE_SORT=0 STEPS=10 timeout 600 tools/bench_workloads.sh million_triangles stress_lit >> gpurun_out/r2k_workloads.log 2>&1
cat gpurun_out/r2k_tests.log gpurun_out/r2k_workloads.log
