set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/pytest_r2c.log 2>&1; tail -8 gpurun_out/pytest_r2c.log
for v in 17 18; do
  CTGB_DOT_VARIANT=$v python scripts/gpu_profile_slice.py complex128 30 > gpurun_out/prof_c128_dot$v.log 2>&1; grep -E "^slice" gpurun_out/prof_c128_dot$v.log; grep -E "K=33554432|N=32 K=2\^25|var=1[78]" gpurun_out/prof_c128_dot$v.log | head -3; head -4 gpurun_out/nodes_complex128_w30.csv
done
python scripts/gpu_profile_slice.py complex64 30 > gpurun_out/prof_c64_fused3.log 2>&1; grep -E "^slice|fusion" gpurun_out/prof_c64_fused3.log; head -6 gpurun_out/nodes_complex64_w30.csv
python scripts/gpu_strip_timing.py complex128 > gpurun_out/strip_c128.json 2> gpurun_out/strip_c128.err; cat gpurun_out/strip_c128.json; tail -3 gpurun_out/strip_c128.err
python scripts/gpu_strip_timing.py complex64 > gpurun_out/strip_c64.json 2> gpurun_out/strip_c64.err; cat gpurun_out/strip_c64.json; tail -3 gpurun_out/strip_c64.err
python bench.py --config peps8x8 --dtype complex64 --no-cpu > gpurun_out/bench_peps_c64.json 2> gpurun_out/bench_peps_c64.err; tail -c 600 gpurun_out/bench_peps_c64.json; tail -3 gpurun_out/bench_peps_c64.err
python bench.py --config peps8x8 --dtype complex128 --no-cpu > gpurun_out/bench_peps_c128.json 2> gpurun_out/bench_peps_c128.err; tail -c 300 gpurun_out/bench_peps_c128.json
python bench.py --config m10 --dtype complex64 --no-cpu > gpurun_out/bench_m10_c64.json 2> gpurun_out/bench_m10_c64.err; tail -c 300 gpurun_out/bench_m10_c64.json; tail -3 gpurun_out/bench_m10_c64.err
python bench.py --config m10 --dtype complex128 --no-cpu > gpurun_out/bench_m10_c128.json 2> gpurun_out/bench_m10_c128.err; tail -c 300 gpurun_out/bench_m10_c128.json
python bench.py --config m12 --dtype complex64 --scaling strong --steps 1 --warmup 0 --no-cpu > gpurun_out/bench_m12_n1.json 2> gpurun_out/bench_m12_n1.err; tail -c 600 gpurun_out/bench_m12_n1.json; tail -3 gpurun_out/bench_m12_n1.err
python bench.py > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 800 gpurun_out/bench_r2c.json; tail -5 gpurun_out/bench_r2c.err
