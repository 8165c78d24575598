cat > /tmp/steady.py <<'PY'
import sys, time, numpy as np
sys.path.insert(0,'.')
import bench, torch
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0)
g.load_problem(prob)
g.set_option('kernel', 1); g.set_option('pool_size', 1<<21)
stream = torch.cuda.ExternalStream(g.stream_ptr())
g.run_histories(0, 4000000); g.synchronize()
for n in (20000000, 20000000):
    g.reset_tallies()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t=time.time(); e0.record(stream); g.run_histories(0, n); e1.record(stream); g.synchronize(); dt=time.time()-t
    print(sys.argv[1], 'n',n,'%.3g hist/s wall'%(n/dt), 'gpu ms %.1f'%e0.elapsed_time(e1), flush=True)
PY
for cfg in "5 5 4 5" "6 6 5 6" "6 5 4 6" "8 6 5 8" "5 6 6 6" "4 4 3 4"; do
  set -- $cfg
  OMC_NVCC_FLAGS="-DOMC_MB_MISC=$1 -DOMC_MB_ESIZE=$2 -DOMC_MB_ECH=$3 -DOMC_MB_EBCA=$4" python ompmc_b200/build.py --force > /dev/null
  python /tmp/steady.py "mb_$1_$2_$3_$4" | tail -1
done
python ompmc_b200/build.py --force > /dev/null
