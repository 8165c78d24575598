set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
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
cfgs = [(1<<21,1,16,16),(1<<21,2,16,16),(1<<22,1,16,32),(1<<21,1,8,16)]
if len(sys.argv) > 2: cfgs = cfgs[:1]
for pool, iters, cross, every in cfgs:
    g.set_option('pool_size', pool); g.set_option('electron_iters', iters); g.set_option('max_cross', cross); g.set_option('check_every', every)
    g.reset_tallies()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t=time.time(); e0.record(stream); g.run_histories(0, n); e1.record(stream); g.synchronize(); dt=time.time()-t
    c=g.counters()
    print(sys.argv[3] if len(sys.argv)>3 else '', 'pool',pool,'iters',iters,'cross',cross,'every',every, '%.3g hist/s wall'%(n/dt), 'gpu ms', e0.elapsed_time(e1), 'waves', c['kernel_launches'], flush=True)
PY
python /tmp/steady.py 20000000 "" f32
ncu --set full --clock-control none --import-source on -k regex:wave_kernel -s 600 -c 4 -o gpurun_out/prof_r01_wave_f32 python /tmp/steady.py 20000000 one > gpurun_out/ncu_full4.log 2>&1
OMC_NVCC_FLAGS="-DOMC_WAVE_F32=0" python ompmc_b200/build.py --force > /dev/null
python /tmp/steady.py 20000000 "" f64
OMC_NVCC_FLAGS="-DOMC_WAVE_MINBLOCKS=4" python ompmc_b200/build.py --force > /dev/null
python /tmp/steady.py 20000000 "" f32mb4
