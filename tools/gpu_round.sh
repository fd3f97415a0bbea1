#!/bin/bash
# one measurement pass on the GPU box: bench line, ncu launch list, one full capture of the
# swimming kernel (tag = $1).  Every ncu command follows the same command run without ncu.
tag=$1
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__thread_inst_executed.sum
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_ncu.csv \
  python bench.py > gpurun_out/${tag}_ncu_bench.log 2>&1
cat gpurun_out/${tag}_bench.json
python bench.py --no-cpu-baseline --no-e2e --steps 4 --warmup 3 > gpurun_out/${tag}_short.json 2>&1 &&
ncu --set full --metrics $M --clock-control none --import-source on -k regex:fb_fast_kernel -s 5 -c 1 -f -o gpurun_out/${tag}_fast \
  python bench.py --no-cpu-baseline --no-e2e --steps 4 --warmup 3 > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_ncu_full.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2>/dev/null; cat gpurun_out/${tag}_bench_reference_arm.json
