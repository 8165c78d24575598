set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 2> gpurun_out/bench_n2.err | grep '^{' | tee gpurun_out/bench_n2.json | cut -c1-300
tail -5 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | grep '^{' | cut -c1-300
