#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wrappers.py -q -m gpu -k "counting_cuckoo" > gpurun_out/c47_pytest.log 2>&1
tail -15 gpurun_out/c47_pytest.log
