#!/bin/bash
# one GPU, last seconds of the round's budget: the out-of-core run of zplt_run_param_file against the oracle, then smoke()
mkdir -p gpurun_out
timeout 75 python -m pytest tests/test_gpu_parity.py -q -m gpu -k out_of_core -s --durations=5 > gpurun_out/c26_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c26_pytest.log
grep -v "^$" gpurun_out/c26_pytest.log | tail -25
timeout 30 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/c26_smoke.log
