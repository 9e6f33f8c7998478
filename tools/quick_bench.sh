#!/bin/bash
# usage: tools/quick_bench.sh [env assignments...]  -> one summary line of the default bench
env "$@" python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('$*: value %.3e obj/s  step %.2f ms  grid %.2f ms (frac %.3f)  factor %.2f ms  ll %.2f ms (%.2f TF)  step %.2f TF  e2e %.3e' % (d['value'], d['ms_per_step'], r['ms_per_launch'], r['frac'], r['factor_kernel']['ms_per_launch'], r['ll_kernel']['ms_per_launch'], r['ll_kernel']['achieved'], r['whole_step']['achieved'], d['e2e']['value']))"
