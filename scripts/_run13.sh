set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_r2m.log 2>&1; tail -6 gpurun_out/pytest_r2m.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "dotstream4 or dmma_32x32 or rowstream_long or long_k or tensor_map or single_operand or check_zero or chunks or dmmastream_long" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02_sanitizer_memcheck.log; tail -12 gpurun_out/r02_sanitizer_memcheck.log
