#!/bin/bash
# Short visit to an 8-GPU box: bench at N = 8 (cross-rank parity probe included) and BASELINE configs[3]
# (32768^2, mvec = 5) on 8 / 4 / 2 GPUs.
tag=$1
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 "${@:4}"; }
run 400 8 29531 bench.py --gpus 8 --steps 100 --warmup 5 --e2e-steps 3 > gpurun_out/bench_g8_$tag.json 2> gpurun_out/bench_g8_$tag.err; echo "bench8 rc=$?"
tail -2 gpurun_out/bench_g8_$tag.err; cut -c1-250 gpurun_out/bench_g8_$tag.json
for g in ${EX_GPUS:-8 4 2}; do
  run 300 $g $((29540+g)) tools/example_time_dist.py 32768 10 5 >> gpurun_out/example_32768_$tag.jsonl 2>> gpurun_out/example_32768_$tag.err; echo "example g=$g rc=$?"
done
cat gpurun_out/example_32768_$tag.jsonl | cut -c1-400
