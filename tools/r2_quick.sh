# per-kernel times per mode / variant build (no tests): bash tools/r2_quick.sh <tag> <mode[:variant]> ...
tag=${1:-r2}; shift
for mv in "$@"; do
m=${mv%%:*}; v=""; [ "$mv" != "$m" ] && v=${mv#*:}
echo "== VY_STREAM_MODE=$m variant '$v'" | tee -a gpurun_out/${tag}_variants.log
VYOLO_LIB_VARIANT=$v VY_STREAM_MODE=$m timeout 120 python tools/kernel_times.py ${CFGS:-coco608_b64 vid320_b256} 2>&1 | cut -c1-150 | tee -a gpurun_out/${tag}_variants.log
done
