set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_full_size_particles_vs_oracle 2>&1 | tail -25 ) > gpurun_out/r2s_tests.log 2>&1
STEPS=10 timeout 600 tools/bench_workloads.sh million_triangles particles_200k > gpurun_out/r2s_workloads.log 2>&1
cat gpurun_out/r2s_tests.log gpurun_out/r2s_workloads.log
