bash scripts/gpu_quick.sh
bash scripts/gpu_prof_v4.sh
