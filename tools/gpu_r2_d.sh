# round 2, call D: sharded handles (shards sharing the one device), plug-in in sharded mode, everything else
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.jsonl
( time timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | grep -E "passed|failed|rror|assert|^FAILED|^tests/|^E " | cut -c1-600 | tail -40 ) 2>&1 | tail -44
