set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tests/run_multigpu_parity.py 2>&1 | tail -30 ) > gpurun_out/r2t_mgpu_parity.log 2>&1
grep -c "OK" gpurun_out/r2t_mgpu_parity.log; grep "world=" gpurun_out/r2t_mgpu_parity.log | grep -v OK
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2t_bench_n8.json 2> gpurun_out/r2t_bench_n8.err
tail -3 gpurun_out/r2t_bench_n8.err
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 --workload instanced --copies 1 > gpurun_out/r2t_bench_n8_instanced.json 2> gpurun_out/r2t_bench_n8_instanced.err
tail -3 gpurun_out/r2t_bench_n8_instanced.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2t_bench_n4.json 2> gpurun_out/r2t_bench_n4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2t_bench_n2.json 2> gpurun_out/r2t_bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2t_bench_n2.json","gpurun_out/r2t_bench_n4.json","gpurun_out/r2t_bench_n8.json","gpurun_out/r2t_bench_n8_instanced.json"):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        pc=j.get("parity_check") or {}
        print(f, "value %.4g ms %.4f e2e %.4g (%.2f ms) scaling %s"%(j["value"], j["ms_per_step"], j["e2e"]["value"], j["e2e"]["ms_per_step"], j["scaling"]), "parity", pc.get("lists_bit_exact"), pc.get("colour_max_abs_diff"), pc.get("vertices"))
        print("   per_rank", [(round(r["frame_ms"],3), round(r["of_which_delivery_ms"],3)) for r in j["details"].get("per_rank",[])])
        print("   before_gather", j.get("before_gather",{}).get("value"), "e2e d2h GB/s", j["e2e"].get("d2h_gbs_all_ranks"), "ceiling", j["e2e"].get("d2h_ceiling_gbs_all_ranks"), "tris", j["config"]["triangles_per_frame"])
    except Exception as e: print(f, "FAILED", e)
PY
