set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2i_tests.log 2>&1
mkdir -p build/variants/base && cp tetra3d_b200/lib/libt3dgeo.so build/variants/base/
timeout 600 tools/bench_variants.sh base c256 mb5 mb7 mb8 st2 st4 nosleep > gpurun_out/r2i_variants.log 2>&1
BENCH_ARGS="--workload instanced_shard" timeout 600 tools/bench_variants.sh base c256 mb5 mb7 mb8 st2 st4 nosleep >> gpurun_out/r2i_variants.log 2>&1
BENCH_ARGS="--workload skinned_crowd" timeout 600 tools/bench_variants.sh base >> gpurun_out/r2i_variants.log 2>&1
BENCH_ARGS="--workload stress_lit" timeout 600 tools/bench_variants.sh base c256 >> gpurun_out/r2i_variants.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_vertex -c 2 -o gpurun_out/r2i_vertex_inst python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --workload instanced_shard > gpurun_out/r2i_ncu2.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_vertex -c 4 -o gpurun_out/r2i_vertex_mt python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2i_ncu3.log 2>&1
cat gpurun_out/r2i_tests.log gpurun_out/r2i_variants.log
