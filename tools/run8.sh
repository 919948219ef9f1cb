mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/exp_phases.py --phases 4 --per-phase 4 --slots 45056 > gpurun_out/exp_phases18.txt 2>&1; cat gpurun_out/exp_phases10.txt | tail -20
