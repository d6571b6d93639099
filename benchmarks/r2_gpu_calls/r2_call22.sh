#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -3 gpurun_out/r2_bench_n2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_n2.json") if l.startswith("{")][-1])
print("N=2 headline", d["value"]/1e9, d["ms_per_step"], "parity", d["parity"]["parity"], "e2e", d["e2e"]["value"]/1e9)
print("parts", d["parts"])
PY
