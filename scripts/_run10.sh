set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "long_k or tensor_map" > gpurun_out/pytest_longk.log 2>&1; tail -8 gpurun_out/pytest_longk.log
timeout 900 python bench.py --config m12 --dtype complex64 --scaling strong --steps 1 --warmup 0 --no-cpu > gpurun_out/bench_m12_n1.json 2> gpurun_out/bench_m12_n1.err; tail -c 400 gpurun_out/bench_m12_n1.json; tail -3 gpurun_out/bench_m12_n1.err
timeout 300 python bench.py --config m10s --dtype complex64 --no-cpu > gpurun_out/bench_m10s_c64.json 2> gpurun_out/bench_m10s_c64.err; tail -c 300 gpurun_out/bench_m10s_c64.json; tail -3 gpurun_out/bench_m10s_c64.err
timeout 300 python bench.py --config m10s --dtype complex128 --no-cpu > gpurun_out/bench_m10s_c128.json 2> gpurun_out/bench_m10s_c128.err; tail -c 300 gpurun_out/bench_m10s_c128.json
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -s > gpurun_out/pytest_configs.log 2>&1; tail -12 gpurun_out/pytest_configs.log
timeout 900 python bench.py > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; tail -c 300 gpurun_out/bench_r2i.json; tail -3 gpurun_out/bench_r2i.err
