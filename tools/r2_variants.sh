tag=${1:-r2}
for v in noq nos ""; do
echo "== variant '$v' (VY_STREAM_MODE=v2)" | tee -a gpurun_out/${tag}_variants.log
VYOLO_LIB_VARIANT=$v VY_STREAM_MODE=v2 timeout 90 python tools/kernel_times.py coco608_b64 vid320_b256 2>&1 | cut -c1-100 | tee -a gpurun_out/${tag}_variants.log
done
