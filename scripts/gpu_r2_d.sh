# round-2 measurement batch: sampler + parity tests on the default and the packed build, A/B throughput, ncu of the packed build
set -x
L=$PWD/ompmc_b200
python -m pytest tests/test_gpu_production_samplers.py tests/test_gpu_parity.py -m gpu -q -s 2>&1 | grep -E "PARITY|passed|failed|FAILED|Error|KS" | cut -c1-700 | tail -40
mkdir -p gpurun_out/par_default; mv gpurun_out/parity_*.json gpurun_out/par_default/ 2>/dev/null
OMPMC_B200_LIB=$L/libompmc_b200_sp.so python -m pytest tests/test_gpu_production_samplers.py tests/test_gpu_parity.py tests/test_gpu_wavefront.py -m gpu -q -s 2>&1 | grep -E "PARITY|passed|failed|FAILED|Error|KS" | cut -c1-700 | tail -40
mkdir -p gpurun_out/par_sp; mv gpurun_out/parity_*.json gpurun_out/par_sp/ 2>/dev/null
for v in fat sp sp6; do OMPMC_B200_LIB=$L/libompmc_b200_$v.so python scripts/steady.py $v 40000000 | tail -1; done
python scripts/steady.py slim 40000000 | tail -1
for ps in 19 20 21 22; do OMPMC_B200_LIB=$L/libompmc_b200_sp.so python scripts/steady.py sp_pool$ps 40000000 pool_size=$((1<<ps)) | tail -1; done
OMPMC_B200_LIB=$L/libompmc_b200_sp.so python scripts/steady.py sp_mv2 40000000 max_virtual=2 | tail -1
OMPMC_B200_LIB=$L/libompmc_b200_sp.so python scripts/steady.py sp_ns20 4000000 nsplit=20 | tail -1
OMPMC_B200_LIB=$L/libompmc_b200_sp.so python bench.py --no-cpu-baseline | cut -c1-330
OMPMC_B200_LIB=$L/libompmc_b200_sp.so python bench.py --workload matrad_prostate --steps 3 --warmup 3 | cut -c1-1500
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
export OMPMC_B200_LIB=$L/libompmc_b200_sp.so
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 100 --csv --log-file gpurun_out/r02_launches_wavefront.csv python /tmp/steady2.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"esize_kernel|edo_kernel|misc_kernel" -s 1200 -c 4 -o gpurun_out/prof_r02_sp python /tmp/steady2.py > gpurun_out/ncu_r02.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
