set -x
mkdir -p gpurun_out
python scripts/gpu_strip_timing.py complex128 > gpurun_out/strip_c128.json 2> gpurun_out/strip_c128.err; cat gpurun_out/strip_c128.json; tail -6 gpurun_out/strip_c128.err
python scripts/gpu_strip_timing.py complex64 > gpurun_out/strip_c64.json 2> gpurun_out/strip_c64.err; cat gpurun_out/strip_c64.json; tail -6 gpurun_out/strip_c64.err
python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/pytest_r2g.log 2>&1; tail -5 gpurun_out/pytest_r2g.log
