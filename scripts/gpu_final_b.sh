set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --steps 6 --warmup 3 2> gpurun_out/bench_r1e.err | grep '^{' | tee gpurun_out/bench_r1e.json | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
