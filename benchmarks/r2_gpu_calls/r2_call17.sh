#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -2 gpurun_out/r2_bench_n8.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_n8.json") if l.startswith("{")][-1])
print("N=8 headline", d["value"]/1e9, d["ms_per_step"], "parity", d["parity"]["parity"], {k:(round(v["total_ms"],1)) for k,v in d["roofline"]["kernels"].items()})
print("e2e", d["e2e"]["value"], "config5", d["config5"]["value"], d["config5"]["ms_per_step"], d["config5"]["parity"])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e --no-parity --opt p2p_copy_lanes=2 > gpurun_out/c17_bench_n8_lanes2.json 2> gpurun_out/c17_bench_n8_lanes2.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/c17_bench_n8_lanes2.json") if l.startswith("{")][-1])
    print("N=8 lanes=2 headline", d["value"]/1e9, d["ms_per_step"], {k:(round(v["total_ms"],1)) for k,v in d["roofline"]["kernels"].items()}, "config5", d["config5"]["value"])
except Exception as e: print("failed", e, open("gpurun_out/c17_bench_n8_lanes2.err").read()[-1500:])
PY
