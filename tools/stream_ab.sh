# A/B of the streaming pass: parity tests of the fused path (default mode), then per-kernel times per mode / variant build
# usage: bash tools/stream_ab.sh <tag> <mode[:variant]> ...
tag=${1:-r2}; shift
timeout 400 python -m pytest tests/test_gpu_postproc.py -m gpu -x -q -k "fused or full_size or finalize or reference_execution or empty or graph" 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
tail -3 gpurun_out/${tag}_tests.log
for mv in "$@"; do
m=${mv%%:*}; v=""; [ "$mv" != "$m" ] && v=${mv#*:}
echo "== VY_STREAM_MODE=$m variant '$v'" | tee -a gpurun_out/${tag}_variants.log
VYOLO_LIB_VARIANT=$v VY_STREAM_MODE=$m timeout 120 python tools/kernel_times.py coco608_b64 vid320_b256 stress416_b128 voc416_b1 2>&1 | cut -c1-150 | tee -a gpurun_out/${tag}_variants.log
done
