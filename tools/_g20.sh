run() { echo "== $*"; }
for v in base pf; do
  lib=tools/abtest/lib_$v.so
  echo "== $v"
  timeout 300 python tools/abtest/run_with_lib.py $lib --workload E --ticks 200 --iters 3 2>&1 | grep "iter 2\|rror"
  timeout 300 python tools/abtest/run_with_lib.py $lib --workload C3 --iters 3 2>&1 | grep "iter 2\|rror"
  timeout 300 python tools/abtest/run_with_lib.py $lib --workload D --iters 3 2>&1 | grep "iter 2\|rror"
done
for sl in 4 16 24; do
  echo "== base slack $sl"
  R2P2_WIN_SLACK=$sl timeout 300 python tools/abtest/run_with_lib.py tools/abtest/lib_base.so --workload E --ticks 200 --iters 3 2>&1 | grep "iter 2\|rror"
  R2P2_WIN_SLACK=$sl timeout 300 python tools/abtest/run_with_lib.py tools/abtest/lib_base.so --workload C3 --pop 8192 --iters 3 2>&1 | grep "iter 2\|rror"
done
for b in 128 256; do
  echo "== base block $b"
  R2P2_BLOCK=$b timeout 300 python tools/abtest/run_with_lib.py tools/abtest/lib_base.so --workload C3 --iters 3 2>&1 | grep "iter 2\|rror"
  R2P2_BLOCK=$b timeout 300 python tools/abtest/run_with_lib.py tools/abtest/lib_base.so --workload C3 --pop 8192 --iters 3 2>&1 | grep "iter 2\|rror"
done
