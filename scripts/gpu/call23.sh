set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2c23_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c23_pytest.log
tail -4 gpurun_out/r2c23_pytest.log
timeout 120 python scripts/run_map.py 256 37888 512 f16x3 4 > gpurun_out/r2c23_map.txt 2>&1
timeout 120 python scripts/run_map.py 64 148000 640 f16x3 3 >> gpurun_out/r2c23_map.txt 2>&1
cat gpurun_out/r2c23_map.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c23_bench.json 2> gpurun_out/r2c23_bench.err; echo "bench rc=$?"
