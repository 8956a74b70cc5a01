set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/pytest_r2b.log 2>&1; tail -15 gpurun_out/pytest_r2b.log
python scripts/gpu_profile_slice.py complex128 30 > gpurun_out/prof_c128_fused2.log 2>&1; grep -E "^slice|fusion" gpurun_out/prof_c128_fused2.log; head -12 gpurun_out/nodes_complex128_w30.csv
python scripts/gpu_profile_slice.py complex64 30 > gpurun_out/prof_c64_fused2.log 2>&1; grep -E "^slice|fusion" gpurun_out/prof_c64_fused2.log; head -8 gpurun_out/nodes_complex64_w30.csv
python scripts/gpu_strip_timing.py complex128 > gpurun_out/strip_c128.json 2> gpurun_out/strip_c128.err; cat gpurun_out/strip_c128.json; tail -3 gpurun_out/strip_c128.err
python scripts/gpu_strip_timing.py complex64 > gpurun_out/strip_c64.json 2> gpurun_out/strip_c64.err; cat gpurun_out/strip_c64.json; tail -3 gpurun_out/strip_c64.err
CTGB_RUN_HUGE=1 python -m pytest tests/test_gpu_configs.py -m gpu -q -k config4 -s > gpurun_out/pytest_r2b_huge.log 2>&1; tail -5 gpurun_out/pytest_r2b_huge.log
python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 1500 gpurun_out/bench_r2b.json; tail -5 gpurun_out/bench_r2b.err
