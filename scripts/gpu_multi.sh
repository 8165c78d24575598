# multi-GPU checks that need more than the one device of the driver's test box: run with `gpurun --gpus N -- bash scripts/gpu_multi.sh N`
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -q -rs 2>&1 | tail -15 | tee gpurun_out/r2_multi_tests_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 1500 gpurun_out/r2_bench_n$N.json; tail -3 gpurun_out/r2_bench_n$N.err
