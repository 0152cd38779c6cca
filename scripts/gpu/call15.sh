set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2c15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c15_pytest.log
tail -16 gpurun_out/r2c15_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c15_smoke.txt 2>&1; tail -2 gpurun_out/r2c15_smoke.txt
