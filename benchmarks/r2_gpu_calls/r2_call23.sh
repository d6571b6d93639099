#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu -k "query or config2" > gpurun_out/c23_pytest.log 2>&1
tail -3 gpurun_out/c23_pytest.log
for o in "cms_aggregate=1" "cms_aggregate=0" "cms_hot_cache=0"; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-micro --no-parity --opt $o > gpurun_out/c23_bench_$o.json 2> gpurun_out/c23_bench_$o.err
python - <<PY
import json
d=json.load(open("gpurun_out/c23_bench_$o.json"))
p=d["parts"]
print("$o", "bloom_check", p["bloom_check"]["value"]/1e9, p["bloom_check"]["parity"], "cms add", p["cms"]["add"]["value"]/1e9, p["cms"]["add"]["parity"], "cms check", p["cms"]["check"]["value"]/1e9)
PY
done
