#!/bin/bash
# one GPU, the last seconds: the pageable (threaded, chunked) block store of the out-of-core run against the oracle
mkdir -p gpurun_out
timeout 13 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "out_of_core and pageable" -p no:cacheprovider > gpurun_out/c28_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c28_pytest.log
grep -v "^$" gpurun_out/c28_pytest.log | tail -8
