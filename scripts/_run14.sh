set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q --maxfail=10 -k "forced_stem or dmma_32x32 or big_slice" > gpurun_out/pytest_r2n.log 2>&1; tail -6 gpurun_out/pytest_r2n.log
timeout 600 python scripts/gpu_profile_slice.py complex128 30 > gpurun_out/prof_c128_r2n.log 2>&1; grep -E "^slice" gpurun_out/prof_c128_r2n.log; grep "var=18" gpurun_out/prof_c128_r2n.log
