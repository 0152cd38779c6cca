set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c3_pytest.log
timeout 300 python scripts/run_csm.py 256 2880000 4096 fp64 4 > gpurun_out/r2c3_csm.txt 2>&1
timeout 300 python scripts/run_csm.py 256 2880000 4096 fp32 4 >> gpurun_out/r2c3_csm.txt 2>&1
timeout 300 python scripts/run_csm.py 64 480000 1024 fp64 4 >> gpurun_out/r2c3_csm.txt 2>&1
B2_BENCH_TRACE=2 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2c3_bench_prof.json 2> gpurun_out/r2c3_bench_prof.err; echo "bench rc=$?"
timeout 300 python scripts/run_map.py 256 37888 512 f16x3 3 > gpurun_out/r2c3_map.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:das_map_tc -s 1 -c 1 -o gpurun_out/r2_k3_micro_full -f python scripts/run_map.py 256 37888 512 f16x3 2 > gpurun_out/r2c3_ncu_k3.log 2>&1
tail -5 gpurun_out/r2c3_pytest.log; cat gpurun_out/r2c3_csm.txt gpurun_out/r2c3_map.txt
