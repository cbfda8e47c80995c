# alternate two library builds over bench.py configs and regimes: bash tools/ab_bench.sh  (variants: old, "")
for v in old ""; do echo "== variant '$v'"; for r in R T; do for c in vid320_b256 voc416_b1; do VYOLO_LIB_VARIANT=$v python bench.py --config $c --regime $r --steps 100 --warmup 5 --no-cpu --no-e2e --no-conv --no-other 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('$r', d['config']['workload'][:24], round(d['value']), round(d.get('value_single_stream')), {k[3:-7]: round(v*1e3,1) for k,v in r['kernel_ms_per_step'].items()})
"; done; done; done
for v in old ""; do for r in R T; do VYOLO_LIB_VARIANT=$v VY_DEBUG_LISTS=1 python bench.py --config vid320_b256 --regime $r --steps 2 --warmup 3 --no-cpu --no-e2e --no-conv --no-other 2>&1 | grep -m1 "streamed lists"; done; done
