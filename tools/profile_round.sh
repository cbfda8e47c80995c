# round-2 profile artefacts (one GPU call): ncu --set full of the fused kernels at the other benchmarked configs, of the
# adjacency NMS kernel, of the streaming alternates; compute-sanitizer over the new code paths
tag=${1:-r2}
for cfg in coco608_b64 stress416_b128 vid320_b256; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'vy_decode_stream|vy_decode_sample|vy_nms_finalize' -s 9 -c 3 -o gpurun_out/${tag}_full_${cfg} python bench.py --config $cfg --steps 3 --warmup 3 --streams 1 --no-graph --no-cpu --no-e2e --no-conv --no-other > gpurun_out/${tag}_full_${cfg}.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'lg_adj' -c 1 -o gpurun_out/${tag}_full_lg_adj python tools/stress_time.py 8 > gpurun_out/${tag}_full_lg_adj.log 2>&1
for m in v2 v3; do
VYOLO_LIB_VARIANT=alt VY_STREAM_MODE=$m timeout 300 ncu --set full --clock-control none -k regex:'vy_decode_stream' -s 10 -c 1 -o gpurun_out/${tag}_full_alt_${m} python tools/kernel_times.py coco608_b64 > gpurun_out/${tag}_full_alt_${m}.log 2>&1
done
VYOLO_LIB_VARIANT=alt VY_STREAM_MODE=v2 timeout 120 python tools/kernel_times.py coco608_b64 vid320_b256 2>&1 | cut -c1-200 > gpurun_out/${tag}_alt_times.log
VYOLO_LIB_VARIANT=alt VY_STREAM_MODE=v3 timeout 120 python tools/kernel_times.py coco608_b64 vid320_b256 2>&1 | cut -c1-200 >> gpurun_out/${tag}_alt_times.log
timeout 120 python tools/kernel_times.py coco608_b64 vid320_b256 2>&1 | cut -c1-200 >> gpurun_out/${tag}_alt_times.log
K='fused_variants or fused_rescue or (large_bit_exact and 3000) or large_argument or finalize_branches or known_answers or hierarchical or voc_match'
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_postproc.py -m gpu -x -q -k "$K" > gpurun_out/${tag}_memcheck_postproc.log 2>&1
tail -4 gpurun_out/${tag}_memcheck_postproc.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 10 python -m pytest tests/test_gpu_postproc.py -m gpu -x -q -k "(fused_variants) or (large_bit_exact and 3000 and False) or known_answers" > gpurun_out/${tag}_racecheck_postproc.log 2>&1
tail -6 gpurun_out/${tag}_racecheck_postproc.log
ls -la gpurun_out/${tag}_full_* | head -20
