#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full ncu captures of the hot kernels; writes
# text/CSV summaries into gpurun_out/ (copied to profiles/ afterwards).  Never a bench number.
set -u
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-paths"
# 1. every launch with its device time (default workload cfg3, fft path + near-field section)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_cfg3.csv $B > $OUT/launches_cfg3.out 2>&1
# 2. full captures of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:fft_rows -s 6 -c 2 -f -o $OUT/prof_fold_fft_rows $B --no-nearfield > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_cols -s 6 -c 1 -f -o $OUT/prof_fft_cols $B --no-nearfield > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:nearfield_kernel -s 2 -c 1 -f -o $OUT/prof_nearfield $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cgemm_tn -s 4 -c 2 -f -o $OUT/prof_cgemm $B --no-nearfield --method dense --workload cfg2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fold_kernel -s 4 -c 1 -f -o $OUT/prof_fold $B --no-nearfield --method fold --workload cfg2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cgemm_tc_kernel -s 8 -c 2 -f -o $OUT/prof_cgemm_tc $B --no-nearfield --method tc --workload cfg2 > /dev/null 2>&1
M='dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|dram__throughput.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|smsp__inst_executed_pipe_fma|sm__pipe_fma_cycles_active|sm__inst_executed_pipe_fp64|l1tex__data_bank_conflicts_pipe_lsu|lts__t_bytes.sum|sm__pipe_fp64_cycles_active|achieved_occupancy|launch__shared_mem_per_block|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum|sm__pipe_tensor.*cycles_active|sm__inst_executed_pipe_tensor|lts__t_sectors_op_read.sum|sm__pipe_tensor_subpipe'
for f in fold_fft_rows fft_cols nearfield cgemm fold cgemm_tc; do
  [ -f $OUT/prof_$f.ncu-rep ] && ncu -i $OUT/prof_$f.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_pick.py "$M" > $OUT/summary_$f.txt
done
ls -la $OUT
