set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scripts/steady.py base
OMC_NVCC_FLAGS="-DOMC_MB_ESIZE=8" python ompmc_b200/build.py --force > /dev/null; python scripts/steady.py esize8
OMC_NVCC_FLAGS="-DOMC_MB_ESIZE=7" python ompmc_b200/build.py --force > /dev/null; python scripts/steady.py esize7
OMC_NVCC_FLAGS="-DOMC_MB_ESIZE=5" python ompmc_b200/build.py --force > /dev/null; python scripts/steady.py esize5
python ompmc_b200/build.py --force > /dev/null
