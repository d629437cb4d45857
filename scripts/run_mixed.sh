python -m pytest tests/ -x -q -m gpu > gpurun_out/t_all.log 2>&1; tail -5 gpurun_out/t_all.log
ROWS_ENGINE=2 COLS_ENGINE=1 COLS_STRIP_MB=0 python scripts/allbins_kernels.py 450 675 1350 2025 3375 3600 4050 > gpurun_out/allbins_mixed3.txt 2>&1; cat gpurun_out/allbins_mixed3.txt
