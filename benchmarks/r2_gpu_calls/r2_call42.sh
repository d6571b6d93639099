#!/bin/bash
# final confirmation of the round: whole GPU suite, smoke, default bench line, wrappers bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/c42_pytest.log 2>&1
tail -4 gpurun_out/c42_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -2 gpurun_out/r2_bench_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n1.json"))
print("headline", d["value"]/1e9, d["ms_per_step"], "parity", d["parity"]["parity"], "e2e", d["e2e"]["value"]/1e9, "launches", d["gpu_launches"], d["clocks"])
for k,v in (d["parts"] or {}).items():
    if "error" in v: print(k, "ERROR", v["error"]); continue
    if "value" in v: print(k, v["value"]/1e9, "G/s parity", v["parity"], {kk:vv for kk,vv in v.items() if "_per_s" in kk})
    else:
        for kk,vv in v.items(): print(k, kk, vv["value"]/1e9, "G/s parity", vv["parity"])
PY
timeout 900 python benchmarks/wrappers_bench.py > gpurun_out/r2_wrappers_bench.jsonl 2> gpurun_out/r2_wrappers_bench.err
tail -3 gpurun_out/r2_wrappers_bench.err
cat gpurun_out/r2_wrappers_bench.jsonl
