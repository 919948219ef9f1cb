# the driver's N-rank launch of bench.py (default exchange: exchange-first steps as CUDA graphs, peer copies + hub push), bounded: the
# whole process group is killed after $2 seconds (default 240)
mkdir -p gpurun_out
N=${1:-2}; LIMIT=${2:-240}
setsid python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_multi_${N}gpu.json 2> gpurun_out/bench_multi_${N}gpu.err &
pid=$!
( sleep $LIMIT; kill -9 -- -$pid 2>/dev/null ) &
watcher=$!
wait $pid
kill $watcher 2>/dev/null
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_multi_${N}gpu.json').read().strip().splitlines()[-1])
    print('N', j['n_gpus'], 'value', round(j['value'],1), 'ms', round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value'],1), 'host', round(j['host_issue_ms_per_step'],3), 'graphs', j['cuda_graphs'], 'kernel_ms', j['roofline'].get('kernel_ms_per_rank'), 'parity', j['parity']['rel_diff'], 'bfs', (j.get('bfs') or {}).get('gteps_geomean'))
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_multi_${N}gpu.err').read()[-2500:])
PY
