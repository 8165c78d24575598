set -x
python -m pytest tests/test_gpu_wavefront.py tests/test_gpu_multi.py tests/test_gpu_edge.py tests/test_gpu_host.py tests/test_matrad.py tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -6
python scripts/small_batches.py 2>&1 | tail -6
python scripts/steady.py ring 40000000 | tail -1
python bench.py --no-cpu-baseline | cut -c1-300
