set -x
timeout 600 python -m pytest tests/test_gpu_wavefront.py -x -q 2>&1 | tail -15
timeout 300 python scripts/handover_probe.py 2>&1 | tail -12
