set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_r1c.err | grep '^{' | tee gpurun_out/bench_r1c.json | cut -c1-400
python scripts/steady.py p2 40000000 pool_size=2097152 | tail -1
python scripts/steady.py p4 40000000 pool_size=4194304 | tail -1
python scripts/steady.py p8 40000000 pool_size=8388608 | tail -1
python scripts/steady.py p4x 80000000 pool_size=4194304 | tail -1
