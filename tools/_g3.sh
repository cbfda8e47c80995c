set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/t3_tests.log
tail -15 gpurun_out/t3_tests.log
timeout 300 python tools/quick_time.py > gpurun_out/t3_quick.log 2>&1
cat gpurun_out/t3_quick.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/t3_bench.json 2> gpurun_out/t3_bench.err
cat gpurun_out/t3_bench.json; tail -5 gpurun_out/t3_bench.err
