#!/bin/bash
# device-timed bench line of every library variant under build/variants/ (value, ms per launch)
# usage: tools/variant_bench.sh <tag> [bench args]
tag=$1; shift
out=gpurun_out/${tag}_variants.txt; : > $out
cp farms_mujoco_b200/libfarmsb200.so /tmp/_keep.so
for v in build/variants/*.so; do
  cp $v farms_mujoco_b200/libfarmsb200.so
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-other-configs --no-export --no-e2e "$@" 2>/dev/null | python -c "
import sys, json
try:
    j = json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$(basename $v)', '%.4g' % j['value'], '%.3f ms' % j['ms_per_step'], 'kernel %.3f ms' % j['roofline']['kernel_ms'], j['config'].get('fast_warps_per_block'))
except Exception as e:
    print('$(basename $v)', 'FAILED', e)" >> $out
done
cp /tmp/_keep.so farms_mujoco_b200/libfarmsb200.so
cat $out
