#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wrappers.py -q -m gpu > gpurun_out/c38_pytest.log 2>&1
tail -40 gpurun_out/c38_pytest.log
