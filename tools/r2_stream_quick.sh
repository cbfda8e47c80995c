# parity tests of the fused path + per-kernel times (new pass only), at two ring budgets
tag=${1:-r2}
timeout 240 python -m pytest tests/test_gpu_postproc.py -m gpu -x -q -k "fused or full_size or finalize or reference_execution or empty or graph" 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
tail -3 gpurun_out/${tag}_tests.log
for kb in 150 205; do
echo "== VY_S2_SMEM_KB=$kb" | tee -a gpurun_out/${tag}_times.log
VY_S2_SMEM_KB=$kb timeout 90 python tools/kernel_times.py coco608_b64 vid320_b256 stress416_b128 2>&1 | grep -v " T " | tee -a gpurun_out/${tag}_times.log
done
