set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 -k "tc05 or tcgen05 or long_k or tensor_map or pair_cases or trees_golden or sycamore or big_slice or c64 or complex64 or config" > gpurun_out/pytest_r2j.log 2>&1; tail -8 gpurun_out/pytest_r2j.log
timeout 900 python bench.py --config m12 --dtype complex64 --scaling strong --steps 1 --warmup 0 --no-cpu > gpurun_out/bench_m12_n1.json 2> gpurun_out/bench_m12_n1.err; tail -c 400 gpurun_out/bench_m12_n1.json; tail -3 gpurun_out/bench_m12_n1.err
