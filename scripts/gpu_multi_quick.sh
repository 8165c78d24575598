# quick multi-GPU re-check after a change of the batch pipeline: tests + bench at N ranks
N=${1:-2}
set -x
python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 2>/dev/null | cut -c1-400
