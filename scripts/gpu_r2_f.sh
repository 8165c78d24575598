set -x
L=$PWD/ompmc_b200
python scripts/steady.py refill6 40000000 | tail -1
OMPMC_B200_LIB=$L/libompmc_b200_misc5.so python scripts/steady.py refill5 40000000 | tail -1
OMPMC_B200_LIB=$L/libompmc_b200_misc5.so python scripts/steady.py refill5_mv16 40000000 max_virtual=16 | tail -1
OMPMC_B200_LIB=$L/libompmc_b200_misc5.so python scripts/steady.py refill5_mv64 40000000 max_virtual=64 | tail -1
python -m pytest tests/test_gpu_production_samplers.py tests/test_gpu_wavefront.py tests/test_gpu_parity.py -m gpu -q -s 2>&1 | grep -E "PARITY|passed|failed|FAILED|Error|KS|skipped" | cut -c1-330 | tail -30
python scripts/steady.py ns20 4000000 nsplit=20 | tail -1
# DRAM traffic per history: every kernel launch of a complete 2e7-history run, two counters
cat > /tmp/traffic.py <<'PY'
import sys
sys.path.insert(0,'.')
import bench
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0); g.load_problem(prob)
g.set_option('kernel', 1); g.set_option('use_graph', 0); g.set_option('overlap', 0)
g.run_histories(0, 20000000); g.synchronize()
print('counters', g.counters())
PY
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_traffic_all_launches.csv python /tmp/traffic.py > gpurun_out/r02_traffic.log 2>&1
tail -2 gpurun_out/r02_traffic.log; wc -l gpurun_out/r02_traffic_all_launches.csv
