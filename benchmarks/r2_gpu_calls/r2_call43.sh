#!/bin/bash
# N = 2: which knob moves the co-running slowdown of pass 1 (baseline 76.7 ms/step vs 60.0 at N = 1)?
mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 \
      --no-e2e --no-micro --no-cpu-baseline --no-parts "$@" 2> gpurun_out/c43_$tag.err | grep '^{' > gpurun_out/c43_$tag.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c43_$tag.json"))
    k=d["roofline"]["kernels"]
    print("$tag", round(d["value"]/1e9,2), "G keys/s", round(d["ms_per_step"],2), "ms", {n:round(v["avg_launch_ms"],2) for n,v in k.items()}, "parity", d["parity"].get("parity"))
except Exception as e:
    print("$tag failed", e)
PY
}
run base
run lanes1 --opt p2p_copy_lanes=1
run ctas4 --opt bloom_part_ctas_per_sm=4
run chunk26 --chunk-keys 67108864
