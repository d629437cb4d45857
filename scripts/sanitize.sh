#!/bin/bash
# compute-sanitizer passes over small C-ABI tests: memcheck (out-of-bounds / misaligned accesses, API
# errors) and racecheck (shared-memory hazards of the hand-rolled mbarrier rings / exchange buffers).  Summaries ->
# gpurun_out/sanitize_*.txt (copied to profiles/ by hand).  Sizes are small: the tools slow kernels 10-100x.
set -u
SEL_MEM='tests/test_farfield_gpu.py::test_dropin_farfield_from_nearfield tests/test_farfield_gpu.py::test_fields_to_farfield_all_bins tests/test_farfield_gpu.py::test_pipelined_tiles_equal_sequential tests/test_nearfield_gpu.py::test_build_nearfield_matches_reference tests/test_nearfield_gpu.py::test_complex64_device_output_and_big tests/test_slab_gpu.py::test_pushed_allgather_epochs'
SEL_SLAB='tests/test_slab_gpu.py::test_virtual_ranks_equal_single_gpu'
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
    python -m pytest $SEL_MEM -x -q -p no:cacheprovider > gpurun_out/sanitize_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 \
    python -m pytest $SEL_SLAB -x -q -p no:cacheprovider -k "1024-4" >> gpurun_out/sanitize_memcheck.txt 2>&1
echo "memcheck(slab) rc=$?" >> gpurun_out/sanitize_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 \
    python -m pytest tests/test_farfield_gpu.py::test_fields_to_farfield_all_bins tests/test_farfield_gpu.py::test_fft_passes_match_numpy -x -q -p no:cacheprovider > gpurun_out/sanitize_racecheck.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitize_racecheck.txt
tail -5 gpurun_out/sanitize_memcheck.txt; tail -5 gpurun_out/sanitize_racecheck.txt
