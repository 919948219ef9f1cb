mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --extra > gpurun_out/bench_r7.json 2> gpurun_out/bench_r7.err; tail -c 3000 gpurun_out/bench_r7.json; tail -5 gpurun_out/bench_r7.err
timeout 600 python tools/exp_bfs.py 2>&1 | cut -c1-200 | tail -22
cd spla_b200/lib
echo "=== test_cuda_backend 12"; timeout 300 ./test_cuda_backend 12 > /tmp/o.txt 2>&1; echo "rc=$?"; tail -15 /tmp/o.txt
