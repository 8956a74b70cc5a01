set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_r2o.log 2>&1; tail -5 gpurun_out/pytest_r2o.log
timeout 600 python scripts/gpu_profile_slice.py complex128 30 > gpurun_out/prof_c128_r2o.log 2>&1; grep -E "^slice" gpurun_out/prof_c128_r2o.log; head -16 gpurun_out/nodes_complex128_w30.csv
timeout 600 python scripts/gpu_profile_slice.py complex64 30 > gpurun_out/prof_c64_r2o.log 2>&1; grep -E "^slice" gpurun_out/prof_c64_r2o.log
