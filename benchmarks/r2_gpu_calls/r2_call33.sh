#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "cuckoo or smoke" > gpurun_out/c33_pytest.log 2>&1
tail -3 gpurun_out/c33_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/c33_bench.json 2> gpurun_out/c33_bench.err
tail -2 gpurun_out/c33_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/c33_bench.json"))
print("headline", d["value"]/1e9, d["ms_per_step"], "parity", d["parity"]["parity"])
c=d["parts"]["cuckoo"]
print("cuckoo add", c["add"]["value"]/1e9, c["add"]["parity"], [ (x["load"], x["Mkeys_per_s"]) for x in c["add"]["load_curve"]][:12])
print("cuckoo check", c["check"]["value"]/1e9, c["check"]["parity"])
PY
