cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -q -rf -k "batch or particles_200k or instanc or frustum or cabi" 2>&1 | tail -15 ) > gpurun_out/r2z_tests.log 2>&1
cat gpurun_out/r2z_tests.log
T3D_HOST_PROFILE=1 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --workload particles_200k > gpurun_out/r2z_p200k.json 2> gpurun_out/r2z_p200k.err
tail -14 gpurun_out/r2z_p200k.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2z_p200k.json").read().strip().splitlines()[-1])
print("value %.3g ms %.4f e2e %.3g"%(j["value"], j["ms_per_step"], j["e2e"]["value"]), j.get("parity_check",{}).get("lists_bit_exact"), {k:v for k,v in j["e2e"].items() if k not in ("path","ceiling")})
PY
