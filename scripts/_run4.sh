set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/pytest_r2d.log 2>&1; tail -6 gpurun_out/pytest_r2d.log
python scripts/gpu_strip_timing.py complex128 > gpurun_out/strip_c128.json 2> gpurun_out/strip_c128.err; cat gpurun_out/strip_c128.json; tail -3 gpurun_out/strip_c128.err
python scripts/gpu_strip_timing.py complex64 > gpurun_out/strip_c64.json 2> gpurun_out/strip_c64.err; cat gpurun_out/strip_c64.json; tail -3 gpurun_out/strip_c64.err
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_slice_dram_fused.csv python scripts/gpu_slice_dram.py complex128 > gpurun_out/slice_dram_fused.log 2>&1; tail -2 gpurun_out/slice_dram_fused.log
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_slice_dram_nofuse.csv python scripts/gpu_slice_dram.py complex128 --nofuse > gpurun_out/slice_dram_nofuse.log 2>&1; tail -2 gpurun_out/slice_dram_nofuse.log
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_slice_dram_fused_c64.csv python scripts/gpu_slice_dram.py complex64 > gpurun_out/slice_dram_fused_c64.log 2>&1; tail -2 gpurun_out/slice_dram_fused_c64.log
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/r02_top_c128 -f python scripts/gpu_node_bench.py complex128 1 --ncu > gpurun_out/ncu_top_c128.log 2>&1; tail -3 gpurun_out/ncu_top_c128.log
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/r02_dot32_c128 -f python scripts/gpu_node_bench.py complex128 1 --ncu --mnk=32,32,33554432 > gpurun_out/ncu_dot32_c128.log 2>&1; tail -3 gpurun_out/ncu_dot32_c128.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_bench_ncu.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-gpu-lib --no-secondary > gpurun_out/bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/bench_under_ncu.log
python bench.py > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -c 600 gpurun_out/bench_r2d.json; tail -5 gpurun_out/bench_r2d.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2.log 2>&1; tail -2 gpurun_out/smoke_r2.log
