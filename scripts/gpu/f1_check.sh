# eigensolver + f1 beamformers: parity tests and timings
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_heig.py tests/test_gpu_das.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log; tail -5 gpurun_out/r2i_pytest.log
timeout 600 python scripts/bench_f1.py 3 1 > gpurun_out/r2i_f1.json 2> gpurun_out/r2i_f1.err; echo "f1 rc=$?"; tail -3 gpurun_out/r2i_f1.err; cat gpurun_out/r2i_f1.json
