#!/bin/bash
# SPLIT variant of the constrained kernel on ground-contact batches, against the single-warp
# kernel (FARMS_B200_CON_SPLIT=0), environments per group as the engine chooses them.
# usage: tools/consplit_bench.sh <tag>
tag=$1
out=gpurun_out/${tag}_consplit.txt; : > $out
run() {  # model envs split
  FARMS_B200_CON_SPLIT=$3 timeout 300 python bench.py --model $1 --envs-per-gpu $2 --steps 20 --warmup 3 \
    --no-cpu-baseline --no-other-configs --no-export --no-e2e 2>gpurun_out/${tag}_consplit.err | python -c "
import sys, json
try:
    j = json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$1 $2 con_split=$3', '%.4g' % j['value'], '%.3f ms' % j['ms_per_step'], 'envs/group', j['config'].get('fast_envs_per_block'), 'split', j['config'].get('constrained_split'))
except Exception as e:
    print('$1 $2 con_split=$3', 'FAILED', e)" >> $out
}
for cfg in ${CONSPLIT_CFGS:-salamander:512 salamander:2048 salamander:4096 salamander:8192 centipede:1024 centipede:4096 centipede:8192}; do
  for sp in 0 1; do run ${cfg%%:*} ${cfg##*:} $sp; done
done
cat $out
