set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scripts/steady.py cur
python scripts/steady.py cur40 40000000 | tail -1
OMC_NVCC_FLAGS="-DOMC_PREFETCH=0" python ompmc_b200/build.py --force > /dev/null; python scripts/steady.py noprefetch
python ompmc_b200/build.py --force > /dev/null
python scripts/ebeam_check.py 4000000
