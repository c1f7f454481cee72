cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T3D_HOST_PROFILE=1 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --workload particles_200k > gpurun_out/r2y_p200k.json 2> gpurun_out/r2y_p200k.err
tail -60 gpurun_out/r2y_p200k.err
