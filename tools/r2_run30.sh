cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 190 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err
tail -3 gpurun_out/r2i_bench_n1.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2i_bench_n1.json").read().strip().splitlines()[-1])
print("value %.3g ms %.4f e2e %.3g"%(j["value"], j["ms_per_step"], j["e2e"]["value"]), j.get("parity_check",{}).get("lists_bit_exact"), {k:v for k,v in j["e2e"].items() if k in ("ms_per_step","frac_of_d2h_ceiling","one_frame_at_a_time")})
print({k:(round(v["ms"]*1e3,1), round(v.get("frac",0),3)) for k,v in j["roofline"]["stages"].items()}, j["roofline"]["traffic"], j["clocks"])
print("cpu", j.get("cpu_baseline",{}).get("value"))
for k,v in (j.get("extra") or {}).get("workloads",{}).items(): print("   ",k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("device_ms","e2e_ms","e2e_two_frames_in_flight_ms","triangles_per_s","skipped","error")})
PY
