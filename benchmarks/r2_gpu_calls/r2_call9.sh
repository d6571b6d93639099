#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/c9_pytest.log 2>&1
tail -15 gpurun_out/c9_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
tail -5 gpurun_out/c9_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/c9_bench.json"))
print("headline", d["value"]/1e9, d["ms_per_step"], "parity", d["parity"]["parity"])
for k,v in (d["parts"] or {}).items():
    if "error" in v: print(k, "ERROR", v["error"]); continue
    if "value" in v: print(k, v["value"]/1e9, "G/s parity", v["parity"], {kk:vv for kk,vv in v.items() if "_per_s" in kk})
    else:
        for kk,vv in v.items(): print(k, kk, vv["value"]/1e9, "G/s parity", vv["parity"])
PY
