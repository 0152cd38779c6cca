set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c2_pytest.log
B2_BENCH_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2c2_bench_trace.json 2> gpurun_out/r2c2_bench_trace.err; echo "bench rc=$?"
for v in "" fftv8mb2; do
  echo "== variant [$v]" >> gpurun_out/r2c2_csm.txt
  if [ -n "$v" ]; then export B2DSP_LIB=$GRAFT_REPO_ROOT/dsptoolbox_b200/csrc/variants/libb2dsp_$v.so; else unset B2DSP_LIB; fi
  timeout 300 python scripts/run_csm.py 256 2880000 4096 fp64 4 >> gpurun_out/r2c2_csm.txt 2>&1
done
unset B2DSP_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:welch_rfft -s 1 -c 1 -o gpurun_out/r2_k1_full -f python scripts/run_csm.py 256 2880000 4096 fp64 2 > gpurun_out/r2c2_ncu_k1.log 2>&1
tail -5 gpurun_out/r2c2_pytest.log
