set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/probe_small_batch.py 2>&1 | tail -12
timeout 300 python tools/probe_pde.py 2>&1 | tail -12
