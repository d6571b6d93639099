#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/c6_pytest.log 2>&1
tail -25 gpurun_out/c6_pytest.log
for o in "bloom_part_ctas_per_sm=4" "bloom_part_ctas_per_sm=3" "bloom_part_ctas_per_sm=2" "bloom_window_log2_bits=26" "bloom_window_log2_bits=28"; do
timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-micro --opt $o > gpurun_out/c6_bench_$o.json 2> gpurun_out/c6_bench_$o.err
python - <<PY
import json
f="gpurun_out/c6_bench_$o"
try:
    d=json.load(open(f+".json")); print("$o", d["value"]/1e9, "Gkeys/s", d["ms_per_step"], {k:v for k,v in d["roofline"]["kernels"].items()})
except Exception as e: print("$o", "failed", e, open(f+".err").read()[-2000:])
PY
done
