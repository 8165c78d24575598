set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -12
python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_r1d.err | grep '^{' | tee gpurun_out/bench_r1d.json | cut -c1-400
tail -3 gpurun_out/bench_r1d.err
python bench.py --steps 5 --warmup 3 --hist-per-step 16777216 --no-cpu-baseline 2>/dev/null | grep '^{' | cut -c1-200
