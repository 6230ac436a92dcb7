#!/bin/bash
bash scripts/diag_ngpu.sh 8 r2i
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2i_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 gpurun_out/r2i_pytest_multi.log | cut -c1-200
