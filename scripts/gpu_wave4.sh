set -x
python -m pytest tests/test_gpu_wavefront.py -x -q 2>&1 | tail -40
cat > /tmp/steady.py <<'PY'
import sys, time, numpy as np
sys.path.insert(0,'.')
import bench, torch
from ompmc_b200.api import GpuTransport
prob, ph, w = bench.build_workload('prostate6mv')
g = GpuTransport(0)
g.load_problem(prob)
g.set_option('kernel', 1)
g.run_histories(0, 200000); g.synchronize()
stream = torch.cuda.ExternalStream(g.stream_ptr())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000000
cfgs = [(1<<20,2,16,4),(1<<21,2,16,4),(1<<22,2,16,4),(1<<21,2,16,16),(1<<21,4,16,16),(1<<21,1,16,16),(1<<20,2,16,4),(1<<21,2,16,4)]
if len(sys.argv) > 2: cfgs = cfgs[:1]
for pool, iters, cross, every in cfgs:
    g.set_option('pool_size', pool); g.set_option('electron_iters', iters); g.set_option('max_cross', cross); g.set_option('check_every', every)
    g.reset_tallies()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t=time.time(); e0.record(stream); g.run_histories(0, n); e1.record(stream); g.synchronize(); dt=time.time()-t
    c=g.counters()
    print('pool',pool,'iters',iters,'cross',cross,'every',every, '%.3g hist/s wall'%(n/dt), 'gpu ms', e0.elapsed_time(e1), 'waves', c['kernel_launches'], flush=True)
PY
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clk.csv &
python /tmp/steady.py 20000000
kill %1
sort gpurun_out/clk.csv | uniq -c | sort -rn | head -8
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 60 --csv --log-file gpurun_out/launches_r01_wave2.csv python /tmp/steady.py 20000000 one > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wave_kernel -s 600 -c 1 -o gpurun_out/prof_r01_wave2 python /tmp/steady.py 20000000 one > gpurun_out/ncu_full3.log 2>&1
