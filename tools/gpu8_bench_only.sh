#!/bin/bash
# bench.py at N = 8 and N = 4 on an 8-GPU box (the driver's scaling run in miniature: probes included).
tag=$1
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 "${@:4}"; }
for g in 8 4; do
  run 300 $g $((29550+g)) bench.py --gpus $g --steps 50 --warmup 5 --e2e-steps 2 > gpurun_out/bench_g${g}_$tag.json 2> gpurun_out/bench_g${g}_$tag.err; echo "bench g=$g rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_g${g}_$tag.json") if l.startswith("{")][-1])
    print("g=$g", d.get("value"), d.get("ms_per_step"), "probe", d.get("parity_probe",{}).get("ok"), "example", d.get("example_probe"), "e2e", d.get("e2e",{}).get("value"), (d.get("e2e",{}).get("pageable") or {}).get("value"), d.get("error"))
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/bench_g${g}_$tag.err").read()[-1500:])
PY
done
