# last GPU call of round 1: new tests first (drop-in user code, device .3ddose writer), then the writer probe + ncu of the
# formatting kernel, then the rest of the GPU suite and a bench line.  Every step bounded; outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 280 python -m pytest tests/test_format.py tests/test_gpu_dropin.py tests/test_gpu_host.py tests/test_tables.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r1c_newtests.log
timeout 150 python scripts/writer_probe.py 2 2000000 2>&1 | tail -3 | tee gpurun_out/r1c_writer.log
timeout 100 ncu --set full --clock-control none -k regex:format_kernel -c 2 -o gpurun_out/prof_r01_format ompmc_b200/host/omc_dosxyz_b200 -p /tmp/wp/p.blob -n 400000 -b 4 -o /tmp/wp/ncu > gpurun_out/r1c_ncu.log 2>&1
timeout 400 python -m pytest tests -m gpu -q --deselect tests/test_format.py --deselect tests/test_gpu_dropin.py --deselect tests/test_gpu_host.py 2>&1 | tail -8 | tee gpurun_out/r1c_suite.log
timeout 200 python bench.py --steps 4 --warmup 3 2> gpurun_out/r1c_bench.err | grep '^{' | tee gpurun_out/r1c_bench.json | cut -c1-400
