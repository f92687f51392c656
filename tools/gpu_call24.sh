#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -k "slab_decomposition or c5_rank" > gpurun_out/c24_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c24_pytest.log
grep -v "^$" gpurun_out/c24_pytest.log | tail -5
