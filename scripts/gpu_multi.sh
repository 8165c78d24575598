# multi-GPU measurements that need more than the one device of the driver's test box:
#   gpurun --gpus N -- bash scripts/gpu_multi.sh N [1mm]
# tests of the in-library NCCL path, bench.py at N ranks (dosxyz headline + omc_matrad), BASELINE config 5 to a measured 1 % sigma
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "1" ]; then TR="python"; fi
python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -12 | tee gpurun_out/r02_multi_tests_n$N.log
$TR bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
cut -c1-420 gpurun_out/r02_bench_n$N.json; tail -2 gpurun_out/r02_bench_n$N.err
$TR bench.py --gpus $N --workload matrad_prostate --steps 6 --warmup 3 > gpurun_out/r02_bench_matrad_n$N.json 2> gpurun_out/r02_bench_matrad_n$N.err
cut -c1-300 gpurun_out/r02_bench_matrad_n$N.json; tail -2 gpurun_out/r02_bench_matrad_n$N.err
$TR scripts/run_config5.py --voxel-mm 2 2>&1 | tail -1 | cut -c1-600
if [ "$2" = "1mm" ]; then $TR scripts/run_config5.py --voxel-mm 1 2>&1 | tail -1 | cut -c1-600; fi
# config 4 at plan scale through the plain-C driver, ONE process over N GPUs (in-product gather): 320 beamlets x 1e6 histories
python - <<'PY'
import sys, os
sys.path.insert(0, '.')
import bench
from ompmc_b200 import problem as P
w = bench.WORKLOADS["prostate6mv"]
media = P.load_blob(P.golden(w["media"])); ph = w["phantom"]()
bl = P.matrad_beamlets(ph, gantry_deg=(0.0, 72.0, 144.0, 216.0, 288.0), nbix=(8, 8), bixel_cm=0.5)
prob = P.build_problem_matrad(media, ph, bl, ecut=0.7, pcut=0.01, cdfinv=(media["cdfinv1_var_6MV"], media["cdfinv2_var_6MV"]))
os.makedirs("/tmp/c4", exist_ok=True); P.save_blob("/tmp/c4/m.blob", prob)
PY
( time ompmc_b200/host/omc_matrad_b200 -p /tmp/c4/m.blob -n 1000000 -b 10 -t 0.001 -o /tmp/c4/g$N -G $N ) 2>&1 | grep -E "Beamlets computed|GPUs|real|Total execution" | tee gpurun_out/r02_config4_c_driver_n$N.log
