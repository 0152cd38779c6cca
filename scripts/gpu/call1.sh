set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?"
for v in "" fftmb2 fftmb3; do
  echo "== variant [$v]" >> gpurun_out/r2c1_csm.txt
  if [ -n "$v" ]; then export B2DSP_LIB=$GRAFT_REPO_ROOT/dsptoolbox_b200/csrc/variants/libb2dsp_$v.so; else unset B2DSP_LIB; fi
  timeout 300 python scripts/run_csm.py 256 2880000 4096 fp64 5 >> gpurun_out/r2c1_csm.txt 2>&1
done
unset B2DSP_LIB
timeout 600 python bench.py --workload cfg3 --steps 3 --warmup 3 > gpurun_out/r2c1_bench_cfg3.json 2> gpurun_out/r2c1_bench_cfg3.err; echo "bench cfg3 rc=$?"
tail -5 gpurun_out/r2c1_pytest.log
