#!/bin/bash
# Round-2 evidence run under gpurun (one GPU): tests, bench line + reference arm with the driver's flags, ncu launch
# list, full captures of the hot kernels.  Summaries land in gpurun_out/ and are copied to profiles/ by hand.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/ -x -q -m gpu > $OUT/r02_tests.log 2>&1; tail -3 $OUT/r02_tests.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r02_bench_n1.json 2> $OUT/r02_bench_n1.err; tail -2 $OUT/r02_bench_n1.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/r02_bench_ref.json 2> $OUT/r02_bench_ref.err; tail -2 $OUT/r02_bench_ref.err
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-paths --no-cfg4"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r02_launches_cfg3.csv $B > $OUT/r02_launches_cfg3.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_rows_tma -s 6 -c 2 -f -o $OUT/r02_rows $B --no-nearfield > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cols_power -s 6 -c 2 -f -o $OUT/r02_cols_power $B --no-nearfield > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:nearfield_kernel -s 2 -c 1 -f -o $OUT/r02_nearfield python scripts/run_nearfield_once.py 4096 3 > /dev/null 2>&1
M='dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|dram__throughput.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|sm__inst_executed_pipe_fp64|smsp__inst_executed.sum|smsp__issue_active.avg.pct|lts__t_bytes.sum|sm__pipe_fp64_cycles_active|launch__shared_mem_per_block|long_scoreboard_per_issue|l1tex__t_sector_hit_rate'
for f in rows cols_power nearfield; do
  [ -f $OUT/r02_$f.ncu-rep ] && ncu -i $OUT/r02_$f.ncu-rep --page raw --csv 2>/dev/null | python scripts/ncu_pick.py "$M" > $OUT/r02_ncu_full_$f.txt
done
ls -la $OUT | tail -14
