set -x
mkdir -p gpurun_out
timeout 200 python scripts/run_config4.py 8 8 1000000 2>&1 | tail -4 | tee gpurun_out/r1h_config4.log
timeout 200 python -m pytest tests/test_gpu_host.py tests/test_matrad.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r1h_tests.log
