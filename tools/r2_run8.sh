set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2h_tests.log 2>&1
mkdir -p build/variants/base && cp tetra3d_b200/lib/libt3dgeo.so build/variants/base/
timeout 600 tools/bench_variants.sh base st2 st4 nosleep sleep256 mb4 > gpurun_out/r2h_variants.log 2>&1
BENCH_ARGS="--workload instanced_shard" timeout 600 tools/bench_variants.sh base st2 st4 nosleep sleep256 mb4 >> gpurun_out/r2h_variants.log 2>&1
BENCH_ARGS="--workload skinned_crowd" timeout 600 tools/bench_variants.sh base nosleep >> gpurun_out/r2h_variants.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_vertex -c 2 -o gpurun_out/r2h_vertex_inst python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --workload instanced_shard > gpurun_out/r2h_ncu2.log 2>&1
cat gpurun_out/r2h_tests.log gpurun_out/r2h_variants.log
