mkdir -p gpurun_out
timeout 600 python tools/exp_phases.py --phases 4,5 --per-phase 99 --slots 47104,45056 > gpurun_out/exp_phases19.txt 2>&1; grep -v "density\|max rel" gpurun_out/exp_phases19.txt | tail -14
