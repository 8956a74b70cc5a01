set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -10
python -m pytest tests/test_gpu_multigpu.py -m gpu -q -s -k "nccl" > gpurun_out/pytest_multigpu_n2.log 2>&1; tail -12 gpurun_out/pytest_multigpu_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config m12 --dtype complex64 --scaling strong --steps 1 --warmup 0 --no-cpu > gpurun_out/bench_m12_n2.json 2> gpurun_out/bench_m12_n2.err; tail -c 500 gpurun_out/bench_m12_n2.json; tail -3 gpurun_out/bench_m12_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_m20_n2.json 2> gpurun_out/bench_m20_n2.err; tail -c 400 gpurun_out/bench_m20_n2.json; tail -3 gpurun_out/bench_m20_n2.err
