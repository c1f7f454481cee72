set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/run_multigpu_parity.py 2>&1 | tail -40 ) > gpurun_out/r2f_mgpu_parity.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
tail -5 gpurun_out/r2f_bench_n2.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
tail -5 gpurun_out/r2f_bench_n1.err
cat gpurun_out/r2f_mgpu_parity.log
python - <<'PY'
import json
for f in ("gpurun_out/r2f_bench_n2.json","gpurun_out/r2f_bench_n1.json"):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3g ms %.4f e2e %.3g"%(j["value"], j["ms_per_step"], j["e2e"]["value"]), j.get("parity_check"), j["details"].get("per_rank"), j.get("before_gather",{}).get("value"), j["e2e"].get("frac_of_d2h_ceiling"), j["e2e"].get("d2h_ceiling_gbs_all_ranks"))
        for k,v in (j.get("extra") or {}).get("workloads",{}).items(): print("   ",k, {kk:vv for kk,vv in v.items() if kk in ("device_ms","e2e_ms","triangles_per_s","build_s","skipped","error")}, {kk:round(vv.get("frac",0),3) for kk,vv in v.get("stages",{}).items()})
    except Exception as e: print(f, "FAILED", e)
PY
