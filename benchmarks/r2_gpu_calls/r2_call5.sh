#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_counting.py -x -q -m gpu -k "bloom or counting or literal or sequence or batches or saturation or algebra" > gpurun_out/c5_pytest.log 2>&1
tail -25 gpurun_out/c5_pytest.log
for o in "bloom_min_chunks=8" "bloom_min_chunks=16" "bloom_apply_cpw_per_sm=2" "bloom_apply_cpw_per_sm=8"; do
timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-micro --opt $o > gpurun_out/c5_bench_$o.json 2> gpurun_out/c5_bench_$o.err
python - <<PY
import json
f="gpurun_out/c5_bench_$o"
try:
    d=json.load(open(f+".json")); print("$o", d["value"]/1e9, "Gkeys/s", d["ms_per_step"], {k:v for k,v in d["roofline"]["kernels"].items()})
except Exception as e: print("$o", "failed", e, open(f+".err").read()[-2000:])
PY
done
