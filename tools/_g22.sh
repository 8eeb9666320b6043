for v in im im2; do
  lib=tools/abtest/lib_$v.so
  echo "== $v"
  for w in C3 D B64k; do
    timeout 300 python tools/abtest/run_with_lib.py $lib --workload $w --iters 3 2>&1 | grep "iter 2\|rror"
  done
done
