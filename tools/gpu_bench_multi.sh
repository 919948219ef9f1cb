# N GPUs: the driver's scaling launch of bench.py
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_multi_${N}gpu.json 2> gpurun_out/bench_multi_${N}gpu.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_multi_${N}gpu.json').read().strip().splitlines()[-1])
    print('N', j['n_gpus'], 'value', j['value'], 'ms', j['ms_per_step'], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'], 'kernel_ms_per_rank', j['roofline'].get('kernel_ms_per_rank'))
    print('bfs', json.dumps(j.get('bfs'))[:400])
    print('parity', json.dumps(j.get('parity'))[:300])
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_multi_${N}gpu.err').read()[-3000:])
PY
