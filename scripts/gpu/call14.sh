set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_das.py tests/test_gpu_fullsize.py tests/test_gpu_cfg3.py -x -q > gpurun_out/r2c14_pytest_map.log 2>&1; echo "pytest map rc=$?" >> gpurun_out/r2c14_pytest_map.log
tail -3 gpurun_out/r2c14_pytest_map.log
for v in "" nourg a8b33 a9b32 a6b44; do
  echo "== variant [$v]" >> gpurun_out/r2c14_map.txt
  if [ -n "$v" ]; then export B2DSP_LIB=$GRAFT_REPO_ROOT/dsptoolbox_b200/csrc/variants/libb2dsp_$v.so; else unset B2DSP_LIB; fi
  timeout 120 python scripts/run_map.py 256 37888 512 f16x3 4 >> gpurun_out/r2c14_map.txt 2>&1; echo "rc=$?" >> gpurun_out/r2c14_map.txt
  timeout 120 python scripts/run_map.py 64 148000 640 f16x3 3 >> gpurun_out/r2c14_map.txt 2>&1; echo "rc=$?" >> gpurun_out/r2c14_map.txt
done
unset B2DSP_LIB
cat gpurun_out/r2c14_map.txt
