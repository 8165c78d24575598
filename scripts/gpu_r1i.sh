# final validation of round 1: full GPU suite, config 4 (64 beamlets per pass vs the whole share in one pass), bench line
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r1i_suite.log
timeout 60 python scripts/run_config4.py 8 8 1000000 2>&1 | tail -3 | tee gpurun_out/r1i_config4.log
timeout 110 python bench.py --steps 4 --warmup 3 2> gpurun_out/r1i_bench.err | grep '^{' | tee gpurun_out/r1i_bench.json | cut -c1-200
