# N GPUs: exchange variants of the overlapped step.  usage: gpu_scaling_ab2.sh N "ov dma streams reserve" ...
mkdir -p gpurun_out
N=$1; shift
i=0
for cfg in "$@"; do
set -- $cfg; i=$((i+1))
tag=$1$2_s$3_r$4
SPLA_B200_OVERLAP=$1 SPLA_B200_OVERLAP_DMA=$2 SPLA_B200_DMA_STREAMS=$3 SPLA_B200_RESERVE_SMS=$4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$i bench.py --gpus $N --steps 20 --warmup 3 --no-bfs --no-vxm --no-plugin --no-cpu-baseline > gpurun_out/bench_sc2_${N}gpu_$tag.json 2> gpurun_out/bench_sc2_${N}gpu_$tag.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_sc2_${N}gpu_$tag.json').read().strip().splitlines()[-1])
    print('overlap $1 dma $2 streams $3 reserve $4 N', j['n_gpus'], 'value', round(j['value'],1), 'ms', round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value'],1), 'kernel_ms_per_rank', j['roofline'].get('kernel_ms_per_rank'), 'parity', j['parity']['rel_diff'])
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_sc2_${N}gpu_$tag.err').read()[-3000:])
PY
done
