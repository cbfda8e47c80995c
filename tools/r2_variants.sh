tag=${1:-r2}
timeout 240 python -m pytest tests/test_gpu_postproc.py -m gpu -x -q -k "fused or full_size or finalize or reference_execution or empty or graph" 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
tail -3 gpurun_out/${tag}_tests.log
for v in ""; do
echo "== variant '$v'" | tee -a gpurun_out/${tag}_variants.log
VYOLO_LIB_VARIANT=$v timeout 90 python tools/kernel_times.py coco608_b64 vid320_b256 stress416_b128 2>&1 | cut -c1-110 | tee -a gpurun_out/${tag}_variants.log
done
