#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_wrappers.py -q -m gpu -x > gpurun_out/c31_pytest.log 2>&1
tail -5 gpurun_out/c31_pytest.log
timeout 900 python benchmarks/wrappers_bench.py > gpurun_out/r2_wrappers_bench.jsonl 2> gpurun_out/r2_wrappers_bench.err
tail -5 gpurun_out/r2_wrappers_bench.err
cat gpurun_out/r2_wrappers_bench.jsonl
