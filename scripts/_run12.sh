set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_r2k.log 2>&1; tail -8 gpurun_out/pytest_r2k.log
timeout 900 python bench.py --config m12 --dtype complex64 --scaling strong --steps 1 --warmup 0 --no-cpu > gpurun_out/bench_m12_n1.json 2> gpurun_out/bench_m12_n1.err; tail -c 300 gpurun_out/bench_m12_n1.json; tail -3 gpurun_out/bench_m12_n1.err
timeout 600 python bench.py > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err; tail -c 300 gpurun_out/bench_r2k.json; tail -3 gpurun_out/bench_r2k.err
