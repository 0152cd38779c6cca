# final r2 measurements after the K1 rewrite: f1 beamformers, bench lines, ncu launch list, ncu --set full of the M=64 map kernel
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 600 python scripts/bench_f1.py 3 1 > gpurun_out/r2f_f1.json 2> gpurun_out/r2f_f1.err; echo "f1 rc=$?"; tail -3 gpurun_out/r2f_f1.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo "bench ref rc=$?"
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_cfg3.json 2> gpurun_out/r2f_bench_cfg3.err; echo "bench cfg3 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"das_|herk_|welch_|gather_|csm_|auto_|steering|delay|stockham" -c 200 --csv --log-file gpurun_out/r2f_bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r2f_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:das_map_tc -s 1 -c 1 -f -o gpurun_out/r2f_k3_m64_full python scripts/run_map.py 64 800000 640 f16x3 2 > gpurun_out/r2f_ncu_m64.log 2>&1; echo "ncu m64 rc=$?"
ls -la gpurun_out | tail -12
