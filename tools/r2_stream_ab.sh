# stream2 (tile streaming) vs the round-1 unit streaming: parity tests first, then per-kernel times
tag=${1:-r2}
timeout 900 python -m pytest tests/test_gpu_postproc.py -m gpu -x -q -k "fused or full_size or finalize or reference_execution or empty or graph" 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
tail -15 gpurun_out/${tag}_tests.log
for v in new v1; do
  echo "== $v" | tee -a gpurun_out/${tag}_times.log
  if [ $v = v1 ]; then export VY_STREAM_V1=1; else unset VY_STREAM_V1; fi
  timeout 300 python tools/kernel_times.py coco608_b64 vid320_b256 stress416_b128 2>&1 | tee -a gpurun_out/${tag}_times.log
done
unset VY_STREAM_V1
for kb in 110 190; do
  echo "== new, VY_S2_SMEM_KB=$kb" | tee -a gpurun_out/${tag}_times.log
  VY_S2_SMEM_KB=$kb timeout 300 python tools/kernel_times.py coco608_b64 vid320_b256 2>&1 | tee -a gpurun_out/${tag}_times.log
done
