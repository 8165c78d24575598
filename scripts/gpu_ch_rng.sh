set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scripts/steady.py ech6
for mb in 5 4; do
OMC_NVCC_FLAGS="-DOMC_MB_ECH=$mb" python ompmc_b200/build.py --force > /dev/null; python scripts/steady.py ech$mb
done
OMC_NVCC_FLAGS="-DOMC_CH_BLOCK_RNG=0" python ompmc_b200/build.py --force > /dev/null; python scripts/steady.py oldrng
python ompmc_b200/build.py --force > /dev/null
