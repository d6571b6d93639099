#!/bin/bash
mkdir -p gpurun_out
for o in "p2p_copy_lanes=4" "p2p_copy_lanes=1"; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e --no-parity --no-port --opt $o --window-log2 28 > gpurun_out/c14_bench_n8_$o.json 2> gpurun_out/c14_bench_n8_$o.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/c14_bench_n8_$o.json") if l.startswith("{")][-1])
    print("N=8 $o headline", d["value"]/1e9, d["ms_per_step"], {k:(v["launches"], round(v["total_ms"],1)) for k,v in d["roofline"]["kernels"].items()}, "config5", d.get("config5",{}).get("value"))
except Exception as e: print("$o failed", e, open("gpurun_out/c14_bench_n8_$o.err").read()[-1500:])
PY
done
