set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -3 gpurun_out/r2_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "bench ref rc=$?"
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err; echo "bench cfg3 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"das_|herk_|welch_|gather_|csm_|auto_|steering|delay|stockham" -c 200 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:das_map_tc -s 3 -c 1 -f -o gpurun_out/r2_k3_full_cfg5 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 python scripts/run_map.py 256 125000 2048 f16x3 1 > gpurun_out/r2_runmap.txt 2>&1
for G in 1000000 500000 250000 125000; do
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:das_map_tc_kernel -c 1 --csv --log-file gpurun_out/r2_map_traffic_G$G.csv python scripts/run_map.py 256 $G 2048 f16x3 1 > /dev/null 2>&1; echo "traffic $G rc=$?"
done
ls -la gpurun_out
