set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2x_bench_n2.json 2> gpurun_out/r2x_bench_n2.err
tail -5 gpurun_out/r2x_bench_n2.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2x_bench_n2.json").read().strip().splitlines()[-1])
print("value %.3g ms %.4f e2e %.3g"%(j["value"], j["ms_per_step"], j["e2e"]["value"]), j.get("parity_check",{}).get("lists_bit_exact"), "launches", j["gpu_launches"])
print({k:v for k,v in j["e2e"].items() if k not in ("path","ceiling")})
print(j.get("before_gather"))
PY
