mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --extra > gpurun_out/bench_r10.json 2> gpurun_out/bench_r10.err; tail -c 4500 gpurun_out/bench_r10.json; tail -5 gpurun_out/bench_r10.err
(cd spla_b200/lib; echo "=== test_cuda_backend 12"; timeout 300 ./test_cuda_backend 12 > /tmp/o.txt 2>&1; echo "rc=$?"; tail -6 /tmp/o.txt)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r10.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
