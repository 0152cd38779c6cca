set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench N=$N rc=$?"
tail -c 1500 gpurun_out/r2_bench_${N}gpu.json
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2c12_pytest_multi.log 2>&1; tail -3 gpurun_out/r2c12_pytest_multi.log
fi
