#!/bin/bash
# usage: tools/quick_bench.sh  -> prints value / kernel ms for CGP_SMALL_WARPS in 1 2 4
for w in 1 2 4; do
  CGP_SMALL_WARPS=$w python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('warps=$w value %.3e obj/s  step %.2f ms  predict %.2f ms (frac %.3f)  ll %.2f ms  e2e %.3e' % (d['value'], d['ms_per_step'], r['ms_per_launch'], r['frac'], r['ll_kernel']['ms_per_launch'], d['e2e']['value']))"
done
