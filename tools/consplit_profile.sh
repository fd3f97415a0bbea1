#!/bin/bash
# ncu capture of the constrained SPLIT kernel on 4,096 SALAMANDERs on the ground (after the same
# command has run without ncu).  usage: tools/consplit_profile.sh <tag> [model envs]
tag=$1; model=${2:-salamander}; envs=${3:-4096}
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__thread_inst_executed.sum
A="--no-cpu-baseline --no-e2e --no-other-configs --no-export --steps 4 --warmup 3"
python bench.py $A --model $model --envs-per-gpu $envs > gpurun_out/${tag}_csplit_${model}${envs}_short.json 2>/dev/null &&
ncu --set full --metrics $M --clock-control none --import-source on -k regex:fb_fastc_split_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_csplit_${model}${envs} python bench.py $A --model $model --envs-per-gpu $envs > gpurun_out/${tag}_csplit_${model}${envs}_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches_ncu_${model}${envs}.csv python bench.py $A --model $model --envs-per-gpu $envs > /dev/null 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
