cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 40 python -m pytest tests/test_literal_restatement.py tests/test_gpu_pose.py tests/test_gpu_async.py tests/test_gpu_cabi.py -m gpu -q -rf -k "not headline_size" 2>&1 | tail -15 ) > gpurun_out/r2l_tests.log 2>&1
cat gpurun_out/r2l_tests.log
