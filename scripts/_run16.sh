set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2p.json 2> gpurun_out/bench_r2p.err; tail -c 300 gpurun_out/bench_r2p.json; tail -3 gpurun_out/bench_r2p.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err; tail -c 600 gpurun_out/bench_ref_r2.json
timeout 600 python scripts/gpu_strip_timing.py complex128 > gpurun_out/strip_c128.json 2> gpurun_out/strip_c128.err; cat gpurun_out/strip_c128.json
timeout 600 python scripts/gpu_strip_timing.py complex64 > gpurun_out/strip_c64.json 2> gpurun_out/strip_c64.err; cat gpurun_out/strip_c64.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.log 2>&1; tail -2 gpurun_out/smoke_r2.log
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_slice_dram_fused.csv python scripts/gpu_slice_dram.py complex128 > gpurun_out/slice_dram_fused.log 2>&1; tail -1 gpurun_out/slice_dram_fused.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_bench_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-gpu-lib --no-secondary > gpurun_out/bench_under_ncu.log 2>&1; tail -c 200 gpurun_out/bench_under_ncu.log
