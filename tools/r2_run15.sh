set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_bake.py tests/test_gpu_parity.py -m gpu -x -q -k "bake or sort_paths or unsupported" 2>&1 | tail -25 ) > gpurun_out/r2o_tests.log 2>&1
cat gpurun_out/r2o_tests.log
