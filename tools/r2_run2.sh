set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2b_tests.log 2>&1
mkdir -p build/variants/base && cp tetra3d_b200/lib/libt3dgeo.so build/variants/base/
timeout 300 tools/bench_variants.sh base k1f4 st3 > gpurun_out/r2b_variants.log 2>&1
BENCH_ARGS="--workload instanced_shard" timeout 300 tools/bench_variants.sh base k1f4 st3 >> gpurun_out/r2b_variants.log 2>&1
BENCH_ARGS="--workload skinned_crowd" timeout 300 tools/bench_variants.sh base k1f4 st3 >> gpurun_out/r2b_variants.log 2>&1
STEPS=10 timeout 300 tools/bench_workloads.sh million_triangles instanced_shard skinned_crowd stress_lit > gpurun_out/r2b_workloads.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_vertex -c 2 -o gpurun_out/r2b_vertex python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2b_ncu.log 2>&1
cat gpurun_out/r2b_tests.log gpurun_out/r2b_variants.log gpurun_out/r2b_workloads.log
