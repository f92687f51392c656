#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "fused_exchange" > gpurun_out/c20_pytest.log 2>&1; tail -4 gpurun_out/c20_pytest.log
