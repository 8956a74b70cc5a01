set -x
mkdir -p gpurun_out
nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv
(python scripts/gen_big_goldens.py 30 > gpurun_out/big30.log 2>&1 &) 
python -m pytest tests -m gpu -q --maxfail=30 -x -k "not beyond_2_31" > gpurun_out/pytest_r2a.log 2>&1; tail -5 gpurun_out/pytest_r2a.log
python scripts/gpu_profile_slice.py complex128 30 > gpurun_out/prof_c128_fused.log 2>&1; tail -32 gpurun_out/prof_c128_fused.log
python scripts/gpu_profile_slice.py complex128 30 --nofuse > gpurun_out/prof_c128_nofuse.log 2>&1; grep "^slice" gpurun_out/prof_c128_nofuse.log
python scripts/gpu_profile_slice.py complex64 30 > gpurun_out/prof_c64_fused.log 2>&1; grep "^slice" gpurun_out/prof_c64_fused.log
python -m pytest tests/test_gpu_round2.py -m gpu -q -k "beyond_2_31" > gpurun_out/pytest_r2a_big.log 2>&1; tail -5 gpurun_out/pytest_r2a_big.log
# wait for the W=2^30 oracle slice (bounded)
for i in $(seq 1 60); do if grep -q appxB_w30 gpurun_out/big30.log 2>/dev/null; then break; fi; sleep 10; done
cat gpurun_out/big30.log
python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 3000 gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
