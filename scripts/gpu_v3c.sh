set -x
cat > /tmp/steady.py <<'PY'
import sys, time, numpy as np
sys.path.insert(0,'.')
import bench, torch
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0)
g.load_problem(prob)
g.set_option('kernel', 1); g.set_option('use_graph', 0); g.set_option('pool_size', 1<<21)
g.run_histories(0, 20000000); g.synchronize()
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 60 --csv --log-file gpurun_out/launches_r01_wave3.csv python /tmp/steady.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"esize_kernel|edo_kernel|misc_kernel" -s 800 -c 4 -o gpurun_out/prof_r01_wave3 python /tmp/steady.py > gpurun_out/ncu_full5.log 2>&1
python bench.py --steps 5 --warmup 3 --hist-per-step 8388608 2>&1 | grep '^{' | tee gpurun_out/bench_wave3.json
