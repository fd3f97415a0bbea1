#!/bin/bash
# warps-per-block sweep of the SLIM layout (device-timed value only)
out=gpurun_out/wpb_sweep.txt; : > $out
for n in 65536 32768 24576; do
 for w in auto 5 6 7 8; do
  if [ $w = auto ]; then unset FARMS_B200_FAST_WPB; else export FARMS_B200_FAST_WPB=$w; fi
  python bench.py --envs-per-gpu $n --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$n wpb=$w', '%.4g' % j['value'], '%.3f ms' % j['ms_per_step'], j['config'].get('fast_warps_per_block'), j['roofline']['kernel'])" >> $out
 done
done
cat $out
