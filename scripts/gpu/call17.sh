set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
for v in "" nomma nosynth noload skel nommanoload; do
  echo "== variant [$v]" >> gpurun_out/r2c17_map.txt
  if [ -n "$v" ]; then export B2DSP_LIB=$GRAFT_REPO_ROOT/dsptoolbox_b200/csrc/variants/libb2dsp_$v.so; else unset B2DSP_LIB; fi
  timeout 120 python scripts/run_map.py 256 37888 512 f16x3 3 >> gpurun_out/r2c17_map.txt 2>&1; echo "rc=$?" >> gpurun_out/r2c17_map.txt
done
unset B2DSP_LIB
cat gpurun_out/r2c17_map.txt
