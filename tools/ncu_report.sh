#!/bin/bash
# usage: tools/ncu_report.sh <report.ncu-rep> <mangled kernel> <out prefix>
# Writes <prefix>_metrics.txt (key raw metrics) and <prefix>_by_source.txt.
set -e
rep=$1; kern=$2; out=$3
ncu -i $rep --page raw --csv > /tmp/_raw.csv 2>/dev/null
python - "$out" <<'PY'
import csv, sys
rows=list(csv.reader(open('/tmp/_raw.csv')))
hdr, units, vals = rows[0], rows[1], rows[2]
want=['gpu__time_duration.sum','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed.avg.per_cycle_active','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__waves_per_multiprocessor','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__sass_thread_inst_executed_op_fadd_pred_on.sum','smsp__sass_thread_inst_executed_op_fmul_pred_on.sum','smsp__sass_thread_inst_executed_op_ffma_pred_on.sum','smsp__thread_inst_executed.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_sector_hit_rate.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed_pipe_fma.sum','sm__cycles_elapsed.max']
with open(sys.argv[1]+'_metrics.txt','w') as f:
    for h,u,v in zip(hdr,units,vals):
        if h in want:
            f.write(f'{h} [{u}] {v}\n')
print(open(sys.argv[1]+'_metrics.txt').read())
PY
ncu -i $rep --page source --csv > /tmp/_src.csv 2>/dev/null
cuobjdump -xelf all farms_mujoco_b200/libfarmsb200.so > /dev/null 2>&1
nvdisasm -g -c fb_engine.sm_100a.cubin > /tmp/_listing.txt 2>/dev/null
rm -f fb_engine.sm_100a.cubin
python tools/ncu_by_line.py /tmp/_src.csv /tmp/_listing.txt $kern farms_mujoco_b200/csrc > ${out}_by_source.txt
