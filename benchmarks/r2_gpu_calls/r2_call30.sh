#!/bin/bash
mkdir -p gpurun_out
timeout 900 python benchmarks/wrappers_bench.py > gpurun_out/r2_wrappers_bench.jsonl 2> gpurun_out/r2_wrappers_bench.err
tail -5 gpurun_out/r2_wrappers_bench.err
cat gpurun_out/r2_wrappers_bench.jsonl
