set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2j_tests.log 2>&1
STEPS=10 timeout 600 tools/bench_workloads.sh million_triangles instanced_shard skinned_crowd particles_200k particles > gpurun_out/r2j_workloads.log 2>&1
LD_LIBRARY_PATH=$PWD/build/variants/trace:$LD_LIBRARY_PATH timeout 200 python tools/vertex_trace.py million_triangles staged > gpurun_out/r2j_trace.log 2>&1
LD_LIBRARY_PATH=$PWD/build/variants/trace:$LD_LIBRARY_PATH timeout 200 python tools/vertex_trace.py million_triangles chained >> gpurun_out/r2j_trace.log 2>&1
LD_LIBRARY_PATH=$PWD/build/variants/trace:$LD_LIBRARY_PATH timeout 200 python tools/vertex_trace.py instanced_shard staged >> gpurun_out/r2j_trace.log 2>&1
cat gpurun_out/r2j_tests.log gpurun_out/r2j_workloads.log gpurun_out/r2j_trace.log
