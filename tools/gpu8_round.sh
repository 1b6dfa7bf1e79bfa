#!/bin/bash
# One visit to an 8-GPU box: bench at N = 8 with the cross-rank parity probe, the all-GPU 8192^2
# step-by-step parity test, and BASELINE configs[3] (32768^2, mvec = 5) on 2 / 4 / 8 GPUs.
tag=$1
mkdir -p gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 "${@:4}"; }
run 600 8 29531 bench.py --gpus 8 --steps 100 --warmup 5 --e2e-steps 3 > gpurun_out/bench_g8_$tag.json 2> gpurun_out/bench_g8_$tag.err; echo "bench8 rc=$?"
tail -2 gpurun_out/bench_g8_$tag.err; cut -c1-250 gpurun_out/bench_g8_$tag.json
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k all_gpus > gpurun_out/pytest_steps_g8_$tag.log 2>&1; tail -3 gpurun_out/pytest_steps_g8_$tag.log
for g in 8 4 2; do
  run 600 $g $((29540+g)) tools/example_time_dist.py 32768 10 5 >> gpurun_out/example_32768_$tag.jsonl 2>> gpurun_out/example_32768_$tag.err; echo "example g=$g rc=$?"
done
cat gpurun_out/example_32768_$tag.jsonl | cut -c1-400
nvidia-smi topo -m > gpurun_out/topo_g8_$tag.txt 2>&1; lscpu | grep -i "numa\|socket\|^CPU(s)" >> gpurun_out/topo_g8_$tag.txt; free -g >> gpurun_out/topo_g8_$tag.txt
