set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 4 --warmup 3 --hist-per-step 1000000 2>&1 | tail -3 | tee gpurun_out/bench_lockstep.json
python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r01_lockstep.csv python bench.py --steps 2 --warmup 3 --hist-per-step 200000 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/ncu_bench.log
