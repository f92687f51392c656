#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_slab.py --ppd 1024 --ranks 2 2>/dev/null | tee gpurun_out/c7_diag_default.log
timeout 600 python tools/diag_slab.py --ppd 1024 --ranks 2 --opt slab_ring=0 2>/dev/null | tee gpurun_out/c7_diag_noslabring.log
timeout 600 python tools/diag_slab.py --ppd 1024 --ranks 2 --opt yring=0 2>/dev/null | tee gpurun_out/c7_diag_noyring.log
timeout 1200 python -m pytest tests -q -m gpu -k "fused_exchange" > gpurun_out/c7_pytest.log 2>&1; tail -5 gpurun_out/c7_pytest.log
