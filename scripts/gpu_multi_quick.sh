# quick multi-GPU re-check after a change of the batch pipeline: (tests +) bench at N ranks; OMPMC_B200_LIB selects an A/B build
N=${1:-2}
set -x
if [ "$2" != "benchonly" ]; then python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -6; fi
for lib in "" "$PWD/ompmc_b200/libompmc_b200_prering.so" ""; do
if [ -n "$lib" ] && [ ! -f "$lib" ]; then continue; fi
OMPMC_B200_LIB="$lib" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 --e2e-steps 0 > gpurun_out/r02_bench_quick_n${N}.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_quick_n${N}.json')); print('lib=$lib', '%.4g'%d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch_group'])"
done
