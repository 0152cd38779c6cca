# K1/K2 after the spectra-layout change: parity tests of stage 1, stage timings of the build variants
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_fft.py tests/test_gpu_csm.py tests/test_gpu_stft.py tests/test_gpu_transfer_functions.py tests/test_gpu_fullsize.py tests/test_gpu_fft_method.py -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -5 gpurun_out/r2k_pytest.log
V=dsptoolbox_b200/csrc/variants
{ for lib in "" $V/libb2dsp_nofuse.so $V/libb2dsp_nopf.so $V/libb2dsp_plain.so; do echo "== lib [$lib]"; B2DSP_LIB=$lib timeout 300 python scripts/run_csm.py 256 2880000 4096 fp64 4; done; echo "== fp32"; timeout 300 python scripts/run_csm.py 256 2880000 4096 fp32 3; echo "== cfg2"; timeout 300 python scripts/run_csm.py 64 480000 4096 fp64 3; echo "== wl1024"; timeout 300 python scripts/run_csm.py 64 480000 1024 fp64 3; } > gpurun_out/r2k_stage1.txt 2>&1
cat gpurun_out/r2k_stage1.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:welch_rfft -s 1 -c 1 -f -o gpurun_out/r2k_k1_full python scripts/run_csm.py 256 2880000 4096 fp64 2 > gpurun_out/r2k_ncu.log 2>&1; echo "ncu rc=$?"
