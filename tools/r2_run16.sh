set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/run_multigpu_parity.py 2>&1 | tail -40 ) > gpurun_out/r2p_mgpu_parity.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2p_bench_n2.json 2> gpurun_out/r2p_bench_n2.err
tail -5 gpurun_out/r2p_bench_n2.err
T3D_DELIVER_INDEXED=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2p_bench_n2_compact.json 2> gpurun_out/r2p_bench_n2_compact.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --workload instanced_shard --copies 1 > gpurun_out/r2p_bench_n2_inst.json 2> gpurun_out/r2p_bench_n2_inst.err
tail -5 gpurun_out/r2p_bench_n2_inst.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 5 --warmup 3 --workload particles_200k > gpurun_out/r2p_bench_n2_part.json 2> gpurun_out/r2p_bench_n2_part.err
grep "world=" gpurun_out/r2p_mgpu_parity.log
python - <<'PY'
import json
for f in ("gpurun_out/r2p_bench_n2.json","gpurun_out/r2p_bench_n2_compact.json","gpurun_out/r2p_bench_n2_inst.json","gpurun_out/r2p_bench_n2_part.json"):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        pc=j.get("parity_check") or {}
        print(f, "value %.4g ms %.4f e2e %.4g scaling %s"%(j["value"], j["ms_per_step"], j["e2e"]["value"], j["scaling"]), "parity", pc.get("lists_bit_exact"), pc.get("colour_max_abs_diff"), pc.get("vertices"))
        print("   per_rank", [(round(r["frame_ms"],3), round(r["of_which_delivery_ms"],3)) for r in j["details"].get("per_rank",[])])
    except Exception as e: print(f, "FAILED", e)
PY
