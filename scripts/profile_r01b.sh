#!/bin/bash
# Round-1 (second half) evidence run under gpurun: tests, bench line, ncu launch list, full captures of the hot kernels.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/ -x -q -m gpu > $OUT/t_all.log 2>&1; tail -3 $OUT/t_all.log
python bench.py > $OUT/bench_r01b.json 2> $OUT/bench_r01b.err; tail -2 $OUT/bench_r01b.err
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-paths"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_cfg3_b.csv $B > $OUT/launches_cfg3_b.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:fft_rows_tma -s 6 -c 2 -f -o $OUT/prof_rows_b $B --no-nearfield > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:cols_power -s 6 -c 2 -f -o $OUT/prof_cols_power_b $B --no-nearfield > /dev/null 2>&1
ls -la $OUT | tail -12
