#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --durations=4 > gpurun_out/c23_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c23_pytest.log
grep -v "^$" gpurun_out/c23_pytest.log | tail -9
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
