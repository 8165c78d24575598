# 2-GPU box: BASELINE config 4 through the plain-C matRad driver at 1 and 2 GPUs, matRad GPU tests, bench at N=2
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 200 python scripts/run_config4.py 8 8 1000000 2>&1 | tail -3 | tee gpurun_out/r1f_config4.log
timeout 200 python -m pytest tests/test_gpu_host.py tests/test_matrad.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r1f_tests.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 2> gpurun_out/r1f_bench2.err | grep '^{' | tee gpurun_out/r1f_bench_n2.json | cut -c1-300
