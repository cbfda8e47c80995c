# A/B of the programmatic-dependent-launch knobs (VY_PDL_MASK: 1 stream, 2 select, 4 finalize; VY_PDL_TRIG: 0 start, 1 end)
for combo in "0 0" "7 0" "7 1" "6 0" "6 1" "4 0" "5 1" "3 1"; do
  set -- $combo
  for c in coco608_b64 voc416_b1 vid320_b256; do
    VY_PDL_MASK=$1 VY_PDL_TRIG=$2 python bench.py --config $c --steps 100 --warmup 5 --no-cpu --no-e2e --no-conv 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('mask=$1 trig=$2', d['config']['workload'][:30], round(d['value']), round(d.get('value_single_stream')), r.get('step_frac_single_stream'))
"
  done
done
