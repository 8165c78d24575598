set -x
PARITY_LOCKSTEP=0 python scripts/parity_full.py tg119_6mv 100000000 gpurun_out/parity_tg119_6mv_v4.json 2>&1 | grep -v "negative ustep" | tail -4
PARITY_LOCKSTEP=0 python scripts/parity_full.py water6mv 100000000 gpurun_out/parity_water6mv_v4.json 2>&1 | grep -v "negative ustep" | tail -4
