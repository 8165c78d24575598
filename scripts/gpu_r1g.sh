set -x
mkdir -p gpurun_out
timeout 200 python scripts/run_config4.py 8 8 1000000 2>&1 | tail -3 | tee gpurun_out/r1g_config4.log
