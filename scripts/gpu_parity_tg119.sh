set -x
PARITY_LOCKSTEP=0 PARITY_GPU_MULT=2 python scripts/parity_full.py tg119_6mv 400000000 gpurun_out/parity_tg119_6mv_v4_4e8.json 2>&1 | grep -v "negative ustep" | tail -4
