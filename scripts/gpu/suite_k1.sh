# full GPU suite + quick bench line + ncu capture of K1 (after the spectra-layout / fused-untangle change)
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -3 gpurun_out/r2m_pytest.log
timeout 300 python scripts/run_csm.py 256 2880000 4096 fp64 4 > gpurun_out/r2m_stage1.txt 2>&1; cat gpurun_out/r2m_stage1.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:welch_rfft -s 1 -c 1 -f -o gpurun_out/r2m_k1_full python scripts/run_csm.py 256 2880000 4096 fp64 2 > gpurun_out/r2m_ncu.log 2>&1; echo "ncu rc=$?"
