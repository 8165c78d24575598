set -x
bash scripts/gpu_prof_v4.sh
python scripts/scaling_voxels.py 33554432 2>&1 | tail -4
