cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 70 python -m pytest tests -m gpu -q -rf -k "not full_size and not headline_size" 2>&1 | tail -12 ) > gpurun_out/r2k_tests.log 2>&1
cat gpurun_out/r2k_tests.log
python -c "import __graft_entry__ as g; g.smoke()"
