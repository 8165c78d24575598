# round-2 final single-GPU record: the driver's own sequence (GPU suite, smoke, bench) + config 4 / config 5 at N = 1
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; cut -c1-260 gpurun_out/r02_bench_n1.json
python bench.py --impl reference > gpurun_out/r02_bench_reference_n1.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_reference_n1.json
python bench.py --workload matrad_prostate --steps 6 --warmup 3 > gpurun_out/r02_bench_matrad_n1.json 2> gpurun_out/r02_bench_matrad_n1.err; cut -c1-260 gpurun_out/r02_bench_matrad_n1.json
python bench.py --nsplit 20 --hist-per-step 4000000 --no-cpu-baseline > gpurun_out/r02_bench_nsplit20_n1.json 2>/dev/null; cut -c1-260 gpurun_out/r02_bench_nsplit20_n1.json
bash scripts/gpu_multi.sh 1 1mm 2>&1 | grep -E "config5|Beamlets computed|Total execution|real" | cut -c1-500
