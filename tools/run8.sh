mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/exp_phases.py --phases 4,5 --per-phase 4 --slots 45056 > gpurun_out/exp_phases9.txt 2>&1; cat gpurun_out/exp_phases9.txt | tail -60
