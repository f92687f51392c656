#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x -k "fused_exchange or slab_decomposition" > gpurun_out/c12_pytest.log 2>&1; tail -30 gpurun_out/c12_pytest.log
timeout 1200 python tools/diag_slab3.py > gpurun_out/c12_diag.log 2>&1; grep -v "Input sigma\|Final sigma\|Loading power\|Using PLT\|Generating ICs" gpurun_out/c12_diag.log | tail -30
