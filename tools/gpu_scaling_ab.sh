# N GPUs: scaling launch: plain / overlapped with ncclAllGather / overlapped with copy-engine peer copies
mkdir -p gpurun_out
N=${1:-2}
i=0
for cfg in "0 0" "1 0" "1 1"; do
set -- $cfg; i=$((i+1))
SPLA_B200_OVERLAP=$1 SPLA_B200_OVERLAP_DMA=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$i bench.py --gpus $N --steps 20 --warmup 3 --no-bfs --no-vxm > gpurun_out/bench_scaling_${N}gpu_$1$2.json 2> gpurun_out/bench_scaling_${N}gpu_$1$2.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_scaling_${N}gpu_$1$2.json').read().strip().splitlines()[-1])
    print('overlap $1 dma $2 N', j['n_gpus'], 'value', round(j['value'],1), 'ms', round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value'],1), 'kernel_ms_per_rank', j['roofline'].get('kernel_ms_per_rank'), 'parity', j['parity']['rel_diff'])
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_scaling_${N}gpu_$1$2.err').read()[-3000:])
PY
done
