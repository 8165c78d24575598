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
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_traffic_all_launches.csv python /tmp/traffic.py > gpurun_out/r02_traffic.log 2>&1
tail -1 gpurun_out/r02_traffic.log; wc -l gpurun_out/r02_traffic_all_launches.csv
