# round 2, call E (2 GPUs): the sharded handle and the sharded plug-in over DISTINCT devices (ncclBroadcast + peer copies)
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 900 python -m pytest tests/test_gpu_dist_shards.py tests/test_spla_integration.py -m gpu -q 2>&1 | grep -E "passed|failed|rror|assert|^FAILED|^E " | cut -c1-600 | tail -30 ) 2>&1 | tail -34
SPLA_CUDA_DEVICES=2 ./spla_b200/lib/test_cuda_backend 13 2>&1 | grep -E "accelerator|failed|FAIL" | head
