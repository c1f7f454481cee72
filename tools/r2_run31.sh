cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 100 python -m pytest tests/test_gpu_pose.py -m gpu -q -rf 2>&1 | tail -25 ) > gpurun_out/r2j_tests.log 2>&1
cat gpurun_out/r2j_tests.log
( timeout 60 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_pose.py -m gpu -q -x -k "interrupted or blended_poses_parity" 2>&1 | tail -4 ) > gpurun_out/r2j_memcheck.log 2>&1
cat gpurun_out/r2j_memcheck.log
