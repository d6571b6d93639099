#!/bin/bash
# round-2 GPU call 1: parity of the new pass-1 kernel + A/B against the round-1 kernel + ncu of both
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/c1_pytest.log 2>&1
tail -15 gpurun_out/c1_pytest.log
for v in 4 3; do
  timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-micro --opt bloom_part_version=$v > gpurun_out/c1_bench_v$v.json 2> gpurun_out/c1_bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c1_bench_v$v.json")); print("v$v", d["value"]/1e9, "Gkeys/s", d["ms_per_step"], {k:v for k,v in d["roofline"]["kernels"].items()})
except Exception as e: print("v$v failed", e)
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bloom_part -s 2 -c 1 -o gpurun_out/c1_prof_v4 python bench.py --steps 1 --warmup 1 --keys 250000000 --no-e2e --no-cpu-baseline --no-micro > gpurun_out/c1_ncu_v4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bloom_part -s 2 -c 1 -o gpurun_out/c1_prof_v3 python bench.py --steps 1 --warmup 1 --keys 250000000 --no-e2e --no-cpu-baseline --no-micro --opt bloom_part_version=3 > gpurun_out/c1_ncu_v3.log 2>&1
ls -la gpurun_out
