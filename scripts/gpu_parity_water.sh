set -x
PARITY_LOCKSTEP=0 PARITY_GPU_MULT=2 python scripts/parity_full.py water6mv 300000000 gpurun_out/parity_water6mv_v4_3e8.json 2>&1 | grep -v "negative ustep" | tail -3
