# compute-sanitizer over slices of the GPU parity tests: memcheck and racecheck on the selection / NMS kernels (fused path
# incl. the in-finalize rescue, large path incl. the adjacency kernel and the tiled kernel), memcheck + racecheck on the
# fusion conv / layout / temporal kernels.   usage (under gpurun): bash tools/sanitize.sh <tag>
tag=${1:-run}
K1='fused_variants or fused_rescue or (large_bit_exact and 3000) or large_argument or finalize_branches or known_answers or hierarchical or voc_match'
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_postproc.py -m gpu -x -q -k "$K1" > gpurun_out/${tag}_memcheck_postproc.log 2>&1
tail -3 gpurun_out/${tag}_memcheck_postproc.log
K2='fused_variants or (large_bit_exact and 3000) or (large_bit_exact and 1025) or known_answers or (finalize_branches)'
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 10 python -m pytest tests/test_gpu_postproc.py -m gpu -x -q -k "$K2" > gpurun_out/${tag}_racecheck_postproc.log 2>&1
tail -3 gpurun_out/${tag}_racecheck_postproc.log
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_fusion_conv.py tests/test_gpu_temporal_tail.py -m gpu -x -q > gpurun_out/${tag}_memcheck_conv.log 2>&1
tail -3 gpurun_out/${tag}_memcheck_conv.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 10 python -m pytest tests/test_gpu_fusion_conv.py -m gpu -x -q -k "not benchmarked" > gpurun_out/${tag}_racecheck_conv.log 2>&1
tail -3 gpurun_out/${tag}_racecheck_conv.log
