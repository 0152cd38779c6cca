set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
for v in "" a8b33 a9b32 a6b44 a9b32d3; do
  echo "== variant [$v]" >> gpurun_out/r2c16_map.txt
  if [ -n "$v" ]; then export B2DSP_LIB=$GRAFT_REPO_ROOT/dsptoolbox_b200/csrc/variants/libb2dsp_$v.so; else unset B2DSP_LIB; fi
  timeout 120 python scripts/run_map.py 256 37888 512 f16x3 4 >> gpurun_out/r2c16_map.txt 2>&1; echo "rc=$?" >> gpurun_out/r2c16_map.txt
done
unset B2DSP_LIB
cat gpurun_out/r2c16_map.txt
