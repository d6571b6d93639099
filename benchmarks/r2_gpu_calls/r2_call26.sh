#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wrappers.py -q -m gpu -x > gpurun_out/c26_pytest.log 2>&1
tail -30 gpurun_out/c26_pytest.log
