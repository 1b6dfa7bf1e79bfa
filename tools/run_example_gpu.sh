mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_example.py -m gpu -x -q > gpurun_out/pytest_example.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_example.log
tail -25 gpurun_out/pytest_example.log
timeout 300 python tools/example_time.py 4096 20 5 > gpurun_out/example_time.log 2>&1; tail -3 gpurun_out/example_time.log
timeout 300 python tools/example_time.py 1024 20 5 >> gpurun_out/example_time.log 2>&1; tail -1 gpurun_out/example_time.log
