#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c7_smoke.log 2>&1; tail -3 gpurun_out/c7_smoke.log
( time timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err ) 2> gpurun_out/c7_time.log
tail -3 gpurun_out/c7_time.log; tail -5 gpurun_out/c7_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/c7_bench.json"))
print("headline", d["value"]/1e9, d["ms_per_step"], "parity", d["parity"], "e2e", d["e2e"]["value"]/1e9 if d["e2e"] else None)
print("roof", d["roofline"]["frac"], {k:(round(v["total_ms"],1), round(v.get("achieved_GBps",0))) for k,v in d["roofline"]["kernels"].items()})
print("atomic", d["atomic_roofline"])
for k,v in (d["parts"] or {}).items():
    if "error" in v: print(k, "ERROR", v["error"]); continue
    if "value" in v: print(k, v["value"]/1e9, "G/s parity", v["parity"], {kk:vv for kk,vv in v.items() if kk.endswith("_per_s")})
    else:
        for kk,vv in v.items(): print(k, kk, vv["value"]/1e9, "G/s parity", vv["parity"], vv["roofline"].get("frac"), vv["roofline"].get("frac_of_atomic_ceiling"))
print("cpu", d["cpu_baseline"])
PY
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c7_ref.json 2> gpurun_out/c7_ref.err ) 2>> gpurun_out/c7_time.log
cat gpurun_out/c7_ref.json | head -c 1500; tail -3 gpurun_out/c7_time.log
