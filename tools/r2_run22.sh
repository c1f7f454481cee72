set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -rf --durations=6 2>&1 | tail -60 ) > gpurun_out/r2v_tests.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err
tail -5 gpurun_out/r2v_bench_n1.err
cat gpurun_out/r2v_tests.log
python - <<'PY'
import json
for f in ("gpurun_out/r2v_bench_n1.json",):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3g ms %.4f e2e %.3g"%(j["value"], j["ms_per_step"], j["e2e"]["value"]), j.get("parity_check",{}).get("lists_bit_exact"), "launches", j["gpu_launches"], j["clocks"])
        print({k:v for k,v in j["e2e"].items() if k not in ("path","ceiling")})
        print({k:(round(v["ms"]*1e3,1), round(v.get("frac",0),3)) for k,v in j["roofline"]["stages"].items()})
        print("cpu", j.get("cpu_baseline",{}).get("value"))
        for k,v in (j.get("extra") or {}).get("workloads",{}).items(): print("   ",k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("device_ms","e2e_ms","triangles_per_s","skipped","error")})
    except Exception as e: print(f, "FAILED", e)
PY
