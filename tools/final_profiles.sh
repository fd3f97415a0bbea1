#!/bin/bash
# ncu captures of the other kernels (each after the same command has run without ncu):
# the SPLIT variant at 8,192 environments and the constrained SPLIT kernel on the ground configurations
tag=$1
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__thread_inst_executed.sum
A="--no-cpu-baseline --no-e2e --no-other-configs --no-export --steps 4 --warmup 3"
python bench.py $A --envs-per-gpu 8192 > gpurun_out/${tag}_split8192_short.json 2>/dev/null &&
ncu --set full --metrics $M --clock-control none --import-source on -k regex:fb_fast_split_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_split8192 python bench.py $A --envs-per-gpu 8192 > gpurun_out/${tag}_split8192_ncu.log 2>&1
tools/consplit_profile.sh $tag salamander 4096
[ -n "$WITH_CENTIPEDE" ] && tools/consplit_profile.sh $tag centipede 8192
ls -la gpurun_out/${tag}_*.ncu-rep
