#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "cuckoo or smoke" > gpurun_out/c20_pytest.log 2>&1
tail -4 gpurun_out/c20_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -2 gpurun_out/r2_bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n1.json"))
print("headline", d["value"]/1e9, d["ms_per_step"], "parity", d["parity"]["parity"], "e2e", d["e2e"]["value"]/1e9)
c=d["parts"]["cuckoo"]
print("cuckoo add", c["add"]["value"]/1e9, c["add"]["parity"], "check", c["check"]["value"]/1e9)
for p in c["add"]["load_curve"]: print("  ", p)
PY
