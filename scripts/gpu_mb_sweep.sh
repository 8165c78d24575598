for v in "5 5 6 6" "5 4 6 6" "5 5 5 6" "5 5 4 6" "5 5 6 5" "5 5 6 4" "4 5 6 6" "6 5 6 6" "4 4 4 4" "5 5 5 5"; do
  set -- $v
  OMC_NVCC_FLAGS="-DOMC_MB_MISC=$1 -DOMC_MB_ESIZE=$2 -DOMC_MB_ECH=$3 -DOMC_MB_EBCA=$4" python ompmc_b200/build.py --force > /dev/null
  python scripts/steady.py "mb_$1_$2_$3_$4" 2>&1 | tail -1
done
python ompmc_b200/build.py --force > /dev/null
