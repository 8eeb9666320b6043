echo "== own4 variant (re-walk, 4 own samples)"
for w in "E --ticks 200" "C3 --pop 8192"; do
  timeout 300 python tools/abtest/run_with_lib.py tools/abtest/lib_own4.so --workload $w --iters 3 2>&1 | grep "iter 2\|rror"
done
