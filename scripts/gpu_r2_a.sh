set -x
compute-sanitizer --tool memcheck python scripts/_diag2.py 2>&1 | grep -v "^=========     Host Frame\|^=========         in" | head -120
python scripts/msdist_diag.py 2>&1 | tail -3
python -m pytest tests/test_gpu_wavefront.py tests/test_gpu_multi.py tests/test_gpu_host.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -8
python scripts/steady.py base 40000000 | tail -1
python scripts/steady.py nolook 40000000 lookahead=0 | tail -1
python bench.py --no-cpu-baseline > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; cat gpurun_out/r2_bench1.json | cut -c1-400
