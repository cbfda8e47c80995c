timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 10 python -m pytest tests/test_gpu_postproc.py -m gpu -x -q \
  -k "known_answers or (fused_equals and 416 and 20) or (large_bit_exact and 1025) or (bit_exact_random and 3000 and 400)" > gpurun_out/racecheck_r1.log 2>&1
tail -12 gpurun_out/racecheck_r1.log
for n in 512 1024; do VY_FIN_NT=$n timeout 200 python bench.py --streams 1 --steps 100 --warmup 5 --no-cpu --no-e2e --no-conv 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print($n, round(d['value']), d['roofline']['kernel_ms_per_step'])"; done
