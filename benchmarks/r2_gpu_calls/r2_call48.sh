#!/bin/bash
# last sanity check of the final build at N = 2 (the driver's scaling run uses the same launch line)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 \
    --no-micro --no-cpu-baseline 2> gpurun_out/c48.err | grep '^{' > gpurun_out/c48_bench_n2.json
tail -3 gpurun_out/c48.err
python - <<PY
import json
d=json.load(open("gpurun_out/c48_bench_n2.json"))
print("N=2", round(d["value"]/1e9,2), "G keys/s", round(d["ms_per_step"],2), "ms parity", d["parity"].get("parity"), "e2e", round(d["e2e"]["value"]/1e9,2), "parts", {k:(v.get("parity"), round(v.get("value",0)/1e9,2)) for k,v in (d.get("parts") or {}).items()})
PY
