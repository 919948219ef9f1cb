# N GPUs: variants of the graph-replayed exchange-first step.  usage: gpu_scaling_ab3.sh N "win hubpush reserve graph" ...
mkdir -p gpurun_out
N=$1; shift
i=0
for cfg in "$@"; do
set -- $cfg; i=$((i+1))
tag=$1_h$2_r$3_g$4
SPLA_B200_WIN=$1 SPLA_B200_HUB_PUSH=$2 SPLA_B200_RESERVE_SMS=$3 SPLA_B200_GRAPH=$4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus $N --steps 20 --warmup 3 --no-bfs --no-vxm --no-plugin --no-cpu-baseline > gpurun_out/bench_sc3_${N}gpu_$tag.json 2> gpurun_out/bench_sc3_${N}gpu_$tag.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/bench_sc3_${N}gpu_$tag.json').read().strip().splitlines()[-1])
    print('win $1 hubpush $2 reserve $3 graph $4 N', j['n_gpus'], 'value', round(j['value'],1), 'ms', round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value'],1), 'host', round(j['host_issue_ms_per_step'],3), j['cuda_graphs'], 'kernel_ms', j['roofline'].get('kernel_ms_per_rank')[:2], 'parity', j['parity']['rel_diff'])
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_sc3_${N}gpu_$tag.err').read()[-2500:])
PY
done
