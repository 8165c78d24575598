# round-2 A/B: azimuth as sincos(2 pi u) instead of the box rejection (OMC_AZIMUTH_SINCOS, default off).  Statistical check first
# (wavefront vs lock-step tests), then throughput.
set -x
python scripts/steady.py base 40000000 | tail -1
OMC_NVCC_FLAGS="-DOMC_AZIMUTH_SINCOS=1" python ompmc_b200/build.py --force > /dev/null
python -m pytest tests/test_gpu_wavefront.py tests/test_gpu_samplers.py -m gpu -q 2>&1 | tail -3
python scripts/steady.py sincos 40000000 | tail -1
python scripts/ebeam_check.py 2>&1 | tail -4
python ompmc_b200/build.py --force > /dev/null
