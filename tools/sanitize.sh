# compute-sanitizer over a small slice of the GPU parity tests: memcheck, then racecheck on the selection / NMS kernels
# usage (under gpurun): bash tools/sanitize.sh <tag>
tag=${1:-run}
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_postproc.py -m gpu -x -q \
  -k "known_answers or fused_variants or finalize_branches or (bit_exact_random and 3000) or bbox_iou or yolo_output_block" > gpurun_out/${tag}_memcheck_postproc.log 2>&1
tail -6 gpurun_out/${tag}_memcheck_postproc.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 10 python -m pytest tests/test_gpu_postproc.py -m gpu -x -q \
  -k "known_answers or (fused_equals and 416 and 20) or finalize_branches or (bit_exact_random and 3000 and 400)" > gpurun_out/${tag}_racecheck_postproc.log 2>&1
tail -10 gpurun_out/${tag}_racecheck_postproc.log
