set -x
mkdir -p gpurun_out
timeout 50 python scripts/run_config4.py 8 8 1000000 2>&1 | tail -3 | tee gpurun_out/r1j_config4.log
timeout 60 python -m pytest tests/test_gpu_host.py tests/test_matrad.py -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r1j_tests.log
