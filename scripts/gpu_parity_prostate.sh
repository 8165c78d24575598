set -x
PARITY_LOCKSTEP=0 python scripts/parity_full.py prostate6mv 250000000 gpurun_out/parity_prostate6mv_v4.json 2>&1 | tail -5
