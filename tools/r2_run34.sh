cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 25 python -m pytest tests/test_gpu_async.py tests/test_gpu_cabi.py tests/test_gpu_parity.py -m gpu -q -rf -x -k "not full_size and not headline_size and not scan_chunk" 2>&1 | tail -8 ) > gpurun_out/r2m_tests.log 2>&1
cat gpurun_out/r2m_tests.log
