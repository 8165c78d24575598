set -x
L=$PWD/ompmc_b200
python scripts/steady.py default 40000000 | tail -1
for v in misc5 ch5 es5; do OMPMC_B200_LIB=$L/libompmc_b200_$v.so python scripts/steady.py $v 40000000 | tail -1; done
python scripts/steady.py pool12 40000000 pool_size=12582912 | tail -1
python scripts/steady.py pool16 40000000 pool_size=16777216 | tail -1
python scripts/steady.py every32 40000000 check_every=32 | tail -1
python scripts/steady.py every8 40000000 check_every=8 | tail -1
python scripts/steady.py drain4k 40000000 drain_threshold=4096 | tail -1
python scripts/steady.py drain32k 40000000 drain_threshold=32768 | tail -1
