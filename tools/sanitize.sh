# compute-sanitizer over a small slice of the GPU parity tests (memcheck, then racecheck on the NMS kernels)
out=gpurun_out/sanitize_$1.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_postproc.py -m gpu -x -q \
  -k "known_answers or fused_variants or (bit_exact_random and 3000) or large_argument_variants or bbox_iou or yolo_output_block" > $out 2>&1
tail -15 $out
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_fusion_conv.py -m gpu -x -q -k "matches_oracle or dwconv or roundtrip" > gpurun_out/sanitize_conv_$1.log 2>&1
tail -8 gpurun_out/sanitize_conv_$1.log
