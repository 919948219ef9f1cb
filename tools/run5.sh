timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python tools/exp_bfs.py 2>&1 | cut -c1-200 | tail -22
