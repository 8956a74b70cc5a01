set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "tensor_map" > gpurun_out/pytest_tmap.log 2>&1; tail -15 gpurun_out/pytest_tmap.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_r2h.log 2>&1; tail -6 gpurun_out/pytest_r2h.log
timeout 600 python scripts/gpu_profile_slice.py complex64 30 > gpurun_out/prof_c64_tmap.log 2>&1; grep -E "^slice" gpurun_out/prof_c64_tmap.log; cp gpurun_out/nodes_complex64_w30.csv gpurun_out/nodes_complex64_w30_tmap.csv
CTGB_NO_TENSOR_MAP=1 timeout 600 python scripts/gpu_profile_slice.py complex64 30 > gpurun_out/prof_c64_notmap.log 2>&1; grep -E "^slice" gpurun_out/prof_c64_notmap.log; cp gpurun_out/nodes_complex64_w30.csv gpurun_out/nodes_complex64_w30_notmap.csv
python - <<'PY'
import csv
a=list(csv.DictReader(open('gpurun_out/nodes_complex64_w30_tmap.csv'))); b=list(csv.DictReader(open('gpurun_out/nodes_complex64_w30_notmap.csv')))
ka={(r['M'],r['N'],r['K'],r['variant']):[] for r in a}
for r in a: ka[(r['M'],r['N'],r['K'],r['variant'])].append(float(r['ms']))
kb={}
for r in b: kb.setdefault((r['M'],r['N'],r['K'],r['variant']),[]).append(float(r['ms']))
rows=[]
for k in ka:
    if k in kb and k[3] in ('9','10','11'):
        rows.append((sum(kb[k])-sum(ka[k]),k,sum(ka[k]),sum(kb[k])))
rows.sort(reverse=True)
for d,k,x,y in rows[:10]+rows[-6:]: print('tmap %.2f  notmap %.2f  diff %+.2f'%(x,y,d),k)
print('tc05 total tmap %.2f notmap %.2f'%(sum(r[2] for r in rows),sum(r[3] for r in rows)))
PY
