# round 2, call F: bench with the plug-in leg + profile hooks test; mxv step launch list (baseline for the fixed-cost work)
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_r2f.json').read().strip().splitlines()[-1])
    print('value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'], 'launches', j['gpu_launches'])
    print('plugin', json.dumps(j.get('plugin')))
    print('cpu', json.dumps(j.get('cpu_baseline')))
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_r2f.err').read()[-3000:])
PY
python - <<PY
import sys, torch
sys.path.insert(0, '.')
from spla_b200.backend import Backend
from spla_b200 import graphs, algorithms
be = Backend(0)
n, Ap, Aj = graphs.rmat(20, 16, seed=1, device=be.device)
ones = torch.ones(Aj.numel(), dtype=torch.int32, device=be.device)
torch.cuda.synchronize()
M = be.csr(n, n, Ap.to(torch.int32), Aj, ones)
be.profile(True)
algorithms.bfs(be, M, int(torch.argmax(Ap[1:] - Ap[:-1]).item()))
for k, v in be.profile_dump().items(): print(k, v)
PY
K='regex:mxv_|mask_count|fill_kernel|hub_pack'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/launches_mxv_r2f.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-vxm --no-bfs --no-plugin > gpurun_out/bench_under_ncu_r2f.log 2>&1
tail -2 gpurun_out/bench_under_ncu_r2f.log | cut -c1-300
