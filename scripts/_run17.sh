set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "tc05 or tcgen05 or long_k or tensor_map or pair_cases or trees_golden or sycamore or big_slice or c64 or complex64 or config" > gpurun_out/pytest_r2q.log 2>&1; tail -6 gpurun_out/pytest_r2q.log
timeout 600 python scripts/gpu_profile_slice.py complex64 30 > gpurun_out/prof_c64_r2q.log 2>&1; grep -E "^slice" gpurun_out/prof_c64_r2q.log; head -12 gpurun_out/nodes_complex64_w30.csv
