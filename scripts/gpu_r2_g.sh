set -x
L=$PWD/ompmc_b200
python scripts/steady.py regions 40000000 | tail -1
OMPMC_B200_LIB=$L/libompmc_b200_noreg.so python scripts/steady.py shared_counter 40000000 | tail -1
python scripts/steady.py regions 40000000 | tail -1
python -m pytest tests/test_gpu_wavefront.py tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_matrad.py -m gpu -q 2>&1 | tail -4
python bench.py --no-cpu-baseline | cut -c1-300
