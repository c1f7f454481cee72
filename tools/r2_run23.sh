set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_async.py -m gpu -q -rf 2>&1 | tail -15 ) > gpurun_out/r2w_tests.log 2>&1
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err
tail -5 gpurun_out/r2w_bench_n1.err
T3D_HOST_PROFILE=1 STEPS=4 timeout 200 tools/bench_workloads.sh particles_200k > gpurun_out/r2w_host_profile.log 2>&1
cat gpurun_out/r2w_tests.log
tail -40 gpurun_out/r2w_host_profile.log
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2w_bench_n1.json").read().strip().splitlines()[-1])
print("value %.3g ms %.4f e2e %.3g"%(j["value"], j["ms_per_step"], j["e2e"]["value"]), j.get("parity_check",{}).get("lists_bit_exact"), "launches", j["gpu_launches"])
print({k:v for k,v in j["e2e"].items() if k not in ("path","ceiling")})
PY
