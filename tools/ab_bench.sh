# alternate two library builds over bench.py configs: bash tools/ab_bench.sh  (variants: old, "")
for i in 1 2; do for v in old ""; do echo "== variant '$v'"; for c in coco608_b64 voc416_b1 vid320_b256; do VYOLO_LIB_VARIANT=$v python bench.py --config $c --steps 100 --warmup 5 --no-cpu --no-e2e --no-conv --no-other 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print(d['config']['workload'][:24], round(d['value']), round(d.get('value_single_stream')), {k[3:-7]: round(v*1e3,1) for k,v in r['kernel_ms_per_step'].items()})
"; done; done; done
