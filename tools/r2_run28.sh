set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -rf --durations=6 2>&1 | tail -30 ) > gpurun_out/r2g_tests.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err
tail -5 gpurun_out/r2g_bench_n1.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2g_reference_arm.json 2> gpurun_out/r2g_reference_arm.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_million_triangles_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2g_launches.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_ -c 24 -o gpurun_out/r2g_frame python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2g_ncu.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_async.py tests/test_gpu_pose.py -m gpu -x -q -k "blend or flight or depth_list or one_camera" > gpurun_out/r2g_memcheck.log 2>&1
tail -5 gpurun_out/r2g_memcheck.log
cat gpurun_out/r2g_tests.log gpurun_out/r2g_smoke.log
python - <<'PY'
import json
for f in ("gpurun_out/r2g_bench_n1.json",):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3g ms %.4f e2e %.3g"%(j["value"], j["ms_per_step"], j["e2e"]["value"]), j.get("parity_check",{}).get("lists_bit_exact"), "launches", j["gpu_launches"], j["clocks"])
        print({k:v for k,v in j["e2e"].items() if k not in ("path","ceiling")})
        print({k:(round(v["ms"]*1e3,1), round(v.get("frac",0),3)) for k,v in j["roofline"]["stages"].items()})
        print("cpu", j.get("cpu_baseline",{}).get("value"))
        for k,v in (j.get("extra") or {}).get("workloads",{}).items(): print("   ",k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("device_ms","e2e_ms","e2e_two_frames_in_flight_ms","triangles_per_s","skipped","error")})
    except Exception as e: print(f, "FAILED", e)
PY
