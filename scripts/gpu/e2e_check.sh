# upload path after the workspace-slot change: stage-1 diagnostics from host memory, CSM tests, bench line with both e2e legs, 2-rank gloo/CPU not needed
set -x
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 600 python scripts/diag_e2e.py > gpurun_out/r2q_diag_e2e.txt 2>&1; cut -c1-330 gpurun_out/r2q_diag_e2e.txt
timeout 900 python -m pytest tests/test_gpu_csm.py tests/test_gpu_fullsize.py tests/test_gpu_stft.py tests/test_gpu_transfer_functions.py tests/test_integration_shim.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2q_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2q_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_pageable'], d['parity']['pass'])"
