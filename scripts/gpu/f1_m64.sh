# Orthogonal beamformer on the tensor-core kernel (golden tests + f1 timings), NG / DLAG variants of the 64-microphone map kernel
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_das.py -m gpu -x -q -k "orthogonal or cleansc or mvdr" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; tail -3 gpurun_out/r2g_pytest.log
timeout 600 python scripts/bench_f1.py 3 1 > gpurun_out/r2g_f1.json 2> gpurun_out/r2g_f1.err; echo "f1 rc=$?"; tail -3 gpurun_out/r2g_f1.err; cat gpurun_out/r2g_f1.json
V=dsptoolbox_b200/csrc/variants
{ for lib in "" $V/libb2dsp_dlag3.so $V/libb2dsp_ng2d3.so $V/libb2dsp_ng2d4.so; do echo "== lib [$lib]"; B2DSP_LIB=$lib timeout 300 python scripts/run_map.py 64 800000 640 f16x3 3; B2DSP_LIB=$lib timeout 300 python -m pytest tests/test_gpu_cfg3.py -m gpu -x -q 2>&1 | tail -1; done; } > gpurun_out/r2g_m64_variants.txt 2>&1
cat gpurun_out/r2g_m64_variants.txt
