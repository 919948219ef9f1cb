mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
    print($N, "value", j["value"], "ms", j["ms_per_step"], "kernel_ms", j["roofline"]["kernel_ms_per_rank"], 'frac', j['roofline']['frac'], 'e2e', j['e2e']['value'])
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_${N}gpu.err').read()[-2000:])
PY
