set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 4 --warmup 3 2> gpurun_out/bench_n8.err | grep '^{' | tee gpurun_out/bench_n8.json | cut -c1-300
tail -3 gpurun_out/bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 4 --steps 4 --warmup 3 2> gpurun_out/bench_n4.err | grep '^{' | tee gpurun_out/bench_n4.json | cut -c1-300
