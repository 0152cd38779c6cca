set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_das.py tests/test_gpu_fullsize.py tests/test_gpu_cfg3.py -x -q > gpurun_out/r2c13_pytest_map.log 2>&1; echo "pytest map rc=$?" >> gpurun_out/r2c13_pytest_map.log
tail -3 gpurun_out/r2c13_pytest_map.log
timeout 120 python scripts/run_map.py 256 37888 512 f16x3 4 > gpurun_out/r2c13_map.txt 2>&1
timeout 120 python scripts/run_map.py 64 148000 640 f16x3 3 >> gpurun_out/r2c13_map.txt 2>&1
cat gpurun_out/r2c13_map.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c13_bench.json 2> gpurun_out/r2c13_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c13_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'], d['roofline']['in_run_cublas'])
PY
