#!/bin/bash
# weak (65,536 envs per GPU) and strong (65,536 envs in total) scaling lines at N GPUs
# usage: tools/scale_run.sh <tag> <N>
tag=$1; N=$2
run() { if [ $N = 1 ]; then python bench.py --gpus 1 "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; fi; }
run --no-cpu-baseline --no-other-configs --no-export > gpurun_out/${tag}_weak_n$N.json 2> gpurun_out/${tag}_weak_n$N.err
run --no-cpu-baseline --no-other-configs --no-export --total-envs 65536 > gpurun_out/${tag}_strong_n$N.json 2> gpurun_out/${tag}_strong_n$N.err
python - <<PY
import json
for mode in ('weak', 'strong'):
    try:
        j = json.loads(open('gpurun_out/${tag}_%s_n$N.json' % mode).read().strip().splitlines()[-1])
        print(mode, 'N=$N', 'envs', j['config']['n_envs'], 'value %.4g' % j['value'], 'ms %.3f' % j['ms_per_step'], 'e2e %.4g' % j['e2e']['value'], j['scaling'], j['clocks'])
    except Exception as exc:
        print(mode, 'N=$N FAILED', exc)
PY
