#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c3_pytest.log 2>&1
tail -15 gpurun_out/c3_pytest.log
for mc in 8 16; do
timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-micro --opt bloom_min_chunks=$mc > gpurun_out/c3_bench_mc$mc.json 2> gpurun_out/c3_bench_mc$mc.err
done
timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-micro --opt bloom_overlap=0 > gpurun_out/c3_bench_noov.json 2> gpurun_out/c3_bench_noov.err
python - <<PY
import json
for f in ("c3_bench_mc8","c3_bench_mc16","c3_bench_noov"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"]/1e9, "Gkeys/s", d["ms_per_step"], {k:v for k,v in d["roofline"]["kernels"].items()})
    except Exception as e: print(f, "failed", e, open(f"gpurun_out/{f}.err").read()[-2000:])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bloom_part|bloom_apply" -s 4 -c 2 -o gpurun_out/c3_prof python bench.py --steps 1 --warmup 1 --keys 250000000 --no-e2e --no-cpu-baseline --no-micro > gpurun_out/c3_ncu.log 2>&1
