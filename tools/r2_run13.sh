set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2m_env.log 2>&1; free -g >> gpurun_out/r2m_env.log; nproc >> gpurun_out/r2m_env.log; nvidia-smi topo -m >> gpurun_out/r2m_env.log 2>&1
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bake.py tests/test_gpu_cabi.py -m gpu -x -q -k "sort_paths or bake or unsupported or instances or one_call" 2>&1 | tail -25 ) > gpurun_out/r2m_tests.log 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/run_multigpu_parity.py 2>&1 | tail -40 ) > gpurun_out/r2m_mgpu_parity.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err
tail -5 gpurun_out/r2m_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --workload instanced_shard --copies 1 > gpurun_out/r2m_bench_n2_inst.json 2> gpurun_out/r2m_bench_n2_inst.err
tail -5 gpurun_out/r2m_bench_n2_inst.err
cat gpurun_out/r2m_env.log gpurun_out/r2m_tests.log gpurun_out/r2m_mgpu_parity.log
python - <<'PY'
import json
for f in ("gpurun_out/r2m_bench_n2.json","gpurun_out/r2m_bench_n2_inst.json"):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.3g ms %.4f e2e %.3g scaling %s"%(j["value"], j["ms_per_step"], j["e2e"]["value"], j["scaling"]), j.get("parity_check"), j["details"].get("per_rank"), "before_gather", j.get("before_gather",{}).get("value"), "e2e frac", j["e2e"].get("frac_of_d2h_ceiling"), j["e2e"].get("d2h_ceiling_gbs_all_ranks"))
        print("   stages", {k:round(v["ms"]*1e3,1) for k,v in j["roofline"]["stages"].items()})
    except Exception as e: print(f, "FAILED", e)
PY
