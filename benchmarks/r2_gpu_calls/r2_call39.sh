#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wrappers.py -q -m gpu -k "expanding or rotating" > gpurun_out/c39_pytest.log 2>&1
tail -3 gpurun_out/c39_pytest.log
timeout 900 python benchmarks/wrappers_bench.py --parts expanding_bloom,rotating_bloom 2>&1 | tail -3
