#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bloom" > gpurun_out/c2_pytest.log 2>&1
tail -5 gpurun_out/c2_pytest.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-micro > gpurun_out/c2_bench_v4.json 2> gpurun_out/c2_bench_v4.err
timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-micro --opt bloom_overlap=0 > gpurun_out/c2_bench_v4_noov.json 2> gpurun_out/c2_bench_v4_noov.err
python - <<PY
import json
for f in ("c2_bench_v4","c2_bench_v4_noov"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"]/1e9, "Gkeys/s", d["ms_per_step"], {k:v for k,v in d["roofline"]["kernels"].items()})
    except Exception as e: print(f, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bloom_part -s 2 -c 1 -o gpurun_out/c2_prof_v4 python bench.py --steps 1 --warmup 1 --keys 250000000 --no-e2e --no-cpu-baseline --no-micro > gpurun_out/c2_ncu_v4.log 2>&1
