set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2a_tests.log 2>&1
mkdir -p build/variants/base && cp tetra3d_b200/lib/libt3dgeo.so build/variants/base/
tools/bench_variants.sh base k1f5 k1f4 > gpurun_out/r2a_variants.log 2>&1
BENCH_ARGS="--workload instanced_shard" tools/bench_variants.sh base k1f5 k1f4 >> gpurun_out/r2a_variants.log 2>&1
STEPS=10 tools/bench_workloads.sh million_triangles instanced_shard skinned_crowd stress_lit > gpurun_out/r2a_workloads.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_vertex -c 2 -o gpurun_out/r2a_vertex python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2a_ncu.log 2>&1
cat gpurun_out/r2a_tests.log gpurun_out/r2a_variants.log gpurun_out/r2a_workloads.log
