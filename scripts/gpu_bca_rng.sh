set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scripts/steady.py ebca6
for mb in 5 4; do
OMC_NVCC_FLAGS="-DOMC_MB_EBCA=$mb" python ompmc_b200/build.py --force > /dev/null; python scripts/steady.py ebca$mb
done
python ompmc_b200/build.py --force > /dev/null
