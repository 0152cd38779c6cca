# full-size map kernel under the power cap: default build vs try_wait with a suspend-time hint
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
V=dsptoolbox_b200/csrc/variants
for lib in "" $V/libb2dsp_hint.so ""; do
  echo "== lib [$lib]"
  B2DSP_LIB=$lib timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])"
done > gpurun_out/r2h_hint.txt 2>&1
cat gpurun_out/r2h_hint.txt
