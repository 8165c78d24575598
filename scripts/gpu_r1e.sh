# ncu --set full with source of the four wave kernels on the final build (steady state, pool 2 Mi) for per-line analysis
set -x
mkdir -p gpurun_out
cat > /tmp/steady2.py <<'PY'
import sys
sys.path.insert(0,'.')
import bench
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0); g.load_problem(prob)
g.set_option('kernel', 1); g.set_option('use_graph', 0); g.set_option('overlap', 0); g.set_option('pool_size', 1<<21)
g.run_histories(0, 20000000); g.synchronize()
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"esize_kernel|edo_kernel|misc_kernel" -s 800 -c 4 -o gpurun_out/prof_r01_v4g python /tmp/steady2.py > gpurun_out/ncu_full12.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
python scripts/steady.py base 40000000 | tail -1
