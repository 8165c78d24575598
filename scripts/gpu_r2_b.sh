set -x
python scripts/msdist_diag.py 2>&1 | tail -12
sed -i 's/for v in (0, 1, 2):/for v in (2, 1):/; s/, 8192, False/, 50000, False/' scripts/_diag2.py
compute-sanitizer --tool memcheck python scripts/_diag2.py 2>&1 | grep -v "^=========     Host Frame\|^=========         in" | grep -v "^variant" | head -60
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
for v in sincos packed; do OMPMC_B200_LIB=$PWD/ompmc_b200/libompmc_b200_$v.so python scripts/steady.py $v 40000000 | tail -1; done
for mv in 2 4; do python scripts/steady.py mv$mv 40000000 max_virtual=$mv | tail -1; done
