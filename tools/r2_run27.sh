cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "soup or tile_boundar or stress_lit or batch" 2>&1 | tail -4 ) > gpurun_out/r2aa_tests.log 2>&1
cat gpurun_out/r2aa_tests.log
STEPS=20 timeout 300 tools/bench_workloads.sh million_triangles instanced_shard skinned_crowd
