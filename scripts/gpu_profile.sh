set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 6 --warmup 3 2>&1 | grep '^{' | tee gpurun_out/bench_wave.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file gpurun_out/launches_r01_wavefront.csv python bench.py --steps 2 --warmup 3 --hist-per-step 500000 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_bench_wave.log 2>&1
cat > /tmp/prof_drv.py <<'PY'
import sys
sys.path.insert(0,'.')
import bench
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0); g.load_problem(prob); g.set_option('kernel', 1)
g.run_histories(0, 3000000); g.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:electron_step -s 60 -c 2 -o gpurun_out/prof_r01_electron_step python /tmp/prof_drv.py > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:photon_flight -s 60 -c 2 -o gpurun_out/prof_r01_photon_flight python /tmp/prof_drv.py > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
