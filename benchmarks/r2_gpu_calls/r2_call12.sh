#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 benchmarks/sharded_check.py --keys 600000 > gpurun_out/r2_sharded_check_n8.txt 2>&1
tail -4 gpurun_out/r2_sharded_check_n8.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -3 gpurun_out/r2_bench_n8.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_n8.json") if l.startswith("{")][-1])
print("N=8 headline", d["value"]/1e9, d["ms_per_step"], "parity", d["parity"])
print("e2e", d["e2e"])
print("kernels", {k:(round(v["total_ms"],1)) for k,v in d["roofline"]["kernels"].items()})
print("config5", d.get("config5"))
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2_bench_n8_reference.json 2>/dev/null
head -c 600 gpurun_out/r2_bench_n8_reference.json
