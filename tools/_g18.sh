for v in b128 b256 b512; do
  lib=tools/abtest/lib_$v.so
  echo "== $v"
  timeout 300 python tools/abtest/run_with_lib.py $lib --workload E --ticks 200 --iters 3 2>&1 | grep "iter 2\|rror"
  timeout 300 python tools/abtest/run_with_lib.py $lib --workload C3 --pop 8192 --iters 3 2>&1 | grep "iter 2\|rror"
  timeout 300 python tools/abtest/run_with_lib.py $lib --workload C3 --iters 3 2>&1 | grep "iter 2\|rror"
  timeout 300 python tools/abtest/run_with_lib.py $lib --workload B --iters 4 2>&1 | grep "iter 3\|rror"
  timeout 300 python tools/abtest/run_with_lib.py $lib --workload D --iters 3 2>&1 | grep "iter 2\|rror"
done
