set -x
mkdir -p gpurun_out
# complex128 top node: the default 4-DMMA policy and the opt-in 3M policy under ncu, source level
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -o /tmp/top4m -f python scripts/gpu_node_bench.py complex128 1 --ncu > gpurun_out/ncu_top4m.log 2>&1; tail -2 gpurun_out/ncu_top4m.log
ncu -i /tmp/top4m.ncu-rep --page source --csv 2>/dev/null | python scripts/ncu_src_summary.py 45 > gpurun_out/r02_src_top4m.txt
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -o /tmp/top3m -f python scripts/gpu_node_bench.py complex128 1 --ncu --variant=12 > gpurun_out/ncu_top3m.log 2>&1; tail -2 gpurun_out/ncu_top3m.log
ncu -i /tmp/top3m.ncu-rep --page source --csv 2>/dev/null | python scripts/ncu_src_summary.py 60 > gpurun_out/r02_src_top3m.txt
ncu -i /tmp/top3m.ncu-rep --page raw --csv > gpurun_out/r02_top3m_raw.csv 2>/dev/null
# K = 16 node, source level
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -o /tmp/k16 -f python scripts/gpu_node_bench.py complex128 1 --ncu --mnk=8388608,128,16 > gpurun_out/ncu_k16.log 2>&1; tail -2 gpurun_out/ncu_k16.log
ncu -i /tmp/k16.ncu-rep --page source --csv 2>/dev/null | python scripts/ncu_src_summary.py 60 > gpurun_out/r02_src_k16.txt
ncu -i /tmp/k16.ncu-rep --page raw --csv > gpurun_out/r02_k16_raw.csv 2>/dev/null
# 32x32 split-K node, source level
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -o /tmp/d32 -f python scripts/gpu_node_bench.py complex128 1 --ncu --mnk=32,32,33554432 > gpurun_out/ncu_d32.log 2>&1; tail -2 gpurun_out/ncu_d32.log
ncu -i /tmp/d32.ncu-rep --page source --csv 2>/dev/null | python scripts/ncu_src_summary.py 50 > gpurun_out/r02_src_d32.txt
du -sh gpurun_out
