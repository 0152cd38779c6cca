set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2c28_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c28_pytest.log
tail -4 gpurun_out/r2c28_pytest.log
timeout 120 python scripts/run_map.py 256 37888 512 f16x3 3 > gpurun_out/r2c28_map.txt 2>&1
timeout 120 python scripts/run_map.py 64 148000 640 f16x3 3 >> gpurun_out/r2c28_map.txt 2>&1
timeout 120 python scripts/run_map.py 32 148000 640 f16x3 3 >> gpurun_out/r2c28_map.txt 2>&1
cat gpurun_out/r2c28_map.txt
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c28_bench_cfg3.json 2> gpurun_out/r2c28_bench_cfg3.err; echo "bench cfg3 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c28_bench_cfg3.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['parity']['pass'])
PY
