#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --clock-control none"
prof() {
  timeout 900 $NCU --set full --import-source on -k "regex:$2" -s $3 -c $4 -o gpurun_out/r2_prof_$1 python benchmarks/profile_parts.py $1 > gpurun_out/r2_prof_$1.log 2>&1
  ncu -i gpurun_out/r2_prof_$1.ncu-rep --page raw --csv > gpurun_out/r2_prof_$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/r2_prof_$1.ncu-rep --page source --csv > gpurun_out/r2_prof_$1.source.csv 2>/dev/null
  rm -f gpurun_out/r2_prof_$1.ncu-rep
}
prof bloom_query "bloom_probe2|bloom_part4" 0 2
prof cuckoo "cuckoo_claim_fixed16|cuckoo_insert_kernel" 0 4
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
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
