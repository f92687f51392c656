#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -k "general_ppd or golden or errors_are" --durations=5 > gpurun_out/c18_pytest.log 2>&1; grep -v "^$" gpurun_out/c18_pytest.log | tail -25
