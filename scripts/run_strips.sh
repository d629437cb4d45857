python -m pytest tests/test_farfield_gpu.py -x -q -m gpu > gpurun_out/t_ff.log 2>&1; tail -5 gpurun_out/t_ff.log
for mb in 0 16 32 48 64 96; do for fu in always never; do
ROWS_ENGINE=2 COLS_ENGINE=1 R16_OCC=3 COLS_STRIP_MB=$mb FUSE=$fu python scripts/allbins_kernels.py 4096 8192 2>&1 | grep -v "^    fft_rows" ; done; done > gpurun_out/allbins_strips.txt 2>&1
cat gpurun_out/allbins_strips.txt
