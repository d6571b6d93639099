#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "cuckoo or smoke" > gpurun_out/c25_pytest.log 2>&1
tail -3 gpurun_out/c25_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err
tail -2 gpurun_out/c25_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/c25_bench.json"))
print("headline", d["value"]/1e9, d["ms_per_step"], "parity", d["parity"]["parity"], "e2e", d["e2e"]["value"]/1e9)
for k,v in (d["parts"] or {}).items():
    if "error" in v: print(k, "ERROR", v["error"]); continue
    if "value" in v: print(k, v["value"]/1e9, "G/s parity", v["parity"])
    else:
        for kk,vv in v.items(): print(k, kk, vv["value"]/1e9, "G/s parity", vv["parity"], vv["roofline"].get("frac_of_atomic_ceiling"))
PY
