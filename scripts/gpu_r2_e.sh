# round-2: full GPU suite on the new default build, register-cap A/B, ncu capture in steady state (pool 8 Mi), bench
set -x
L=$PWD/ompmc_b200
python -m pytest tests -m gpu -q -s 2>&1 | grep -E "PARITY|passed|failed|FAILED|Error|KS|skipped" | cut -c1-400 | tail -40
mkdir -p gpurun_out/par_r02; mv gpurun_out/parity_*.json gpurun_out/par_r02/ 2>/dev/null
python scripts/steady.py default 40000000 | tail -1
for v in mb7 mb8 mbx; do OMPMC_B200_LIB=$L/libompmc_b200_$v.so python scripts/steady.py $v 40000000 | tail -1; done
python scripts/steady.py ns20 4000000 nsplit=20 | tail -1
python scripts/steady.py ns20_march 4000000 nsplit=20 max_cross=32 | tail -1
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; cut -c1-300 gpurun_out/r02_bench_n1.json
cat > /tmp/steady2.py <<'PY'
import sys
sys.path.insert(0,'.')
import bench
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0); g.load_problem(prob)
g.set_option('kernel', 1); g.set_option('use_graph', 0); g.set_option('overlap', 0); g.set_option('pool_size', 1<<23)
g.run_histories(0, 30000000); g.synchronize()
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 100 --csv --log-file gpurun_out/r02_launches_wavefront.csv python /tmp/steady2.py > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"esize_kernel|edo_kernel|misc_kernel" -s 300 -c 4 -o gpurun_out/prof_r02 python /tmp/steady2.py > gpurun_out/ncu_r02.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
