set -x
bash scripts/gpu_prof_v4.sh
python scripts/steady.py base 40000000 | tail -1
OMC_NVCC_FLAGS="-DOMC_WARP_AGGREGATE_DOSE=1" python ompmc_b200/build.py --force > /dev/null; python scripts/steady.py warpagg 40000000 | tail -1
python ompmc_b200/build.py --force > /dev/null
