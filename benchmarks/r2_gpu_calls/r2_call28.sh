#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wrappers.py -q -m gpu -k "heavy or stream or returns" > gpurun_out/c28_pytest.log 2>&1
tail -40 gpurun_out/c28_pytest.log
