# BASELINE configs 1-3 from input files (own host + the reference's user code drop-in), writer probe on the same tallies
set -x
mkdir -p gpurun_out
df -h /tmp | tail -1; free -g | head -2
timeout 300 python scripts/run_configs.py 1.0 2>&1 | tail -6 | tee gpurun_out/r1d_configs.log
timeout 120 python scripts/writer_probe.py 2 2000000 2>&1 | tail -2 | tee gpurun_out/r1d_writer_f2.log
timeout 300 python scripts/writer_probe.py 3 2000000 2>&1 | tail -2 | tee gpurun_out/r1d_writer_f3.log
